"""The `PokerRL` compat namespace exposes the modules / names that examples/run_cfrp_example.py:9-12,15 imports."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_example_imports_resolve_to_b200_implementation():
    code = ("from PokerRL.cfr.CFRPlus import CFRPlus\n"
            "from PokerRL.game import bet_sets\n"
            "from PokerRL.game.games import DiscretizedNLLeduc\n"
            "from PokerRL.rl.base_cls.workers.ChiefBase import ChiefBase\n"
            "from PokerRL._.CrayonWrapper import CrayonWrapper\n"
            "import pokerrl_b200.cfr.CFRPlus as m\n"
            "assert CFRPlus is m.CFRPlus and bet_sets.POT_ONLY == [1.0] and DiscretizedNLLeduc.BIG_BLIND == 100\n"
            "c = ChiefBase(t_prof=None); w = CrayonWrapper(name='x', path_log_storage=None, chief_handle=c, "
            "runs_distributed=False, runs_cluster=False)\nprint('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "pokerrl_b200", "compat"), ROOT]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_evaluator_modules_resolve():
    """the evaluator side of the plugin surface: PokerRL.eval.{br,lbr}, PokerRange, EvalAgentBase"""
    code = ("from PokerRL.eval.lbr.LocalLBRWorker import LocalLBRWorker\n"
            "from PokerRL.eval.lbr.LBRArgs import LBRArgs\n"
            "from PokerRL.eval.br.LocalBRMaster import LocalBRMaster\n"
            "from PokerRL.game.PokerRange import PokerRange\n"
            "from PokerRL.rl.base_cls.EvalAgentBase import EvalAgentBase\n"
            "import pokerrl_b200.eval.lbr.LocalLBRWorker as m\n"
            "assert LocalLBRWorker is m.LocalLBRWorker and LBRArgs(lbr_check_to_round=1).lbr_check_to_round == 1\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "pokerrl_b200", "compat"), ROOT]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
