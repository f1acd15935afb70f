"""drop-in namespace: PokerRL.eval.lbr"""
