"""Numerical settings the reference takes from gpflow.settings (GPflow 1.5.1 defaults; not part of /root/reference)."""
jitter = 1e-6          # settings.jitter / settings.numerics.jitter_level, used at kernels.py:431,463,578,656; models.py:65
float_type = "float32"  # device arithmetic (the reference runs float64 on TF)
workspace_budget_bytes = 8 << 30  # upper bound for the increment-Gram chunk buffer of one K() call
# differentiable route (autodiff.py): when the float64 static-kernel Gram of one call exceeds this many bytes it is evaluated in
# blocks under activation checkpointing (recomputed in the backward pass: about 8x less memory, about 1.3x the time)
autodiff_gram_budget_bytes = 2 << 30
