"""Arguments of the LBR evaluator (`PokerRL/eval/lbr/LBRArgs.py:11-75`): the bet sizes LBR may use (here: they must be the
agent's, see LocalLBRWorker), the number of hands per seat, and up to which street LBR only check / calls (the paper's
recommendation for four-street games: the turn)."""


class LBRArgs:
    def __init__(self, lbr_bet_set=None, n_lbr_hands_per_seat=30000, lbr_check_to_round=None, n_parallel_lbr_workers=1,
                 use_gpu_for_batch_eval=True, DISTRIBUTED=False):
        self.lbr_bet_set = lbr_bet_set
        self.n_lbr_hands = n_lbr_hands_per_seat
        self.lbr_check_to_round = lbr_check_to_round
        self.n_workers = n_parallel_lbr_workers if DISTRIBUTED else 1
        self.use_gpu_for_batch_eval = use_gpu_for_batch_eval
        self.DISTRIBUTED = DISTRIBUTED
