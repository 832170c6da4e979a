"""tensorflow.contrib stand-in (test infrastructure only; see ../__init__.py)."""
from . import stateless  # noqa: F401
