positive = "positive"


class Logistic:
    def __init__(self, a=0.0, b=1.0):
        self.a, self.b = a, b
