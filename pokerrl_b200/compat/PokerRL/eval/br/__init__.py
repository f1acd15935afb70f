"""drop-in namespace: PokerRL.eval.br"""
