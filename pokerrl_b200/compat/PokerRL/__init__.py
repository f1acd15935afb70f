"""Drop-in `PokerRL` namespace backed by pokerrl_b200: put `pokerrl_b200/compat` on
PYTHONPATH *instead of* the reference and the reference's CFR example scripts run unchanged on the GPU."""
