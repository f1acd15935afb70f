"""`pycrayon` stub (absent in this image): PokerRL/_/CrayonWrapper.py:6,25 only needs the name."""


class CrayonClient:
    def __init__(self, *a, **k):
        raise ConnectionError("pycrayon stub: no crayon server in this sandbox")
