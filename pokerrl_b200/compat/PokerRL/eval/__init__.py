"""drop-in namespace: PokerRL.eval"""
