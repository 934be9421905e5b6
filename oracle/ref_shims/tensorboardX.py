"""Inert stand-in for tensorboardX (logging only; reference mt/mvae/stats.py:22). Test infrastructure."""


class SummaryWriter:

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None
