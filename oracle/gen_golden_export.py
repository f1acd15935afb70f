"""Golden PokerViz tree exports produced by RUNNING THE REFERENCE (TEST INFRASTRUCTURE; build container only).

    python oracle/gen_golden_export.py     # writes tests/golden/export_<game>_<state>.json.gz

`PublicTree.get_tree_as_dict()` (PublicTree.py:143-144, 313-420) of
  state "built":   right after build_tree()  (no strategy / values: "Not Computed")
  state "uniform": after fill_uniform_random() + compute_ev()
"""
import gzip
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden_cfr as gg  # noqa: E402

OUT = gg.OUT


def main():
    for game in ("StandardLeduc", "NLLeduc_POT"):
        tree, _ = gg._make_tree(game)
        for state in ("built", "uniform"):
            if state == "uniform":
                tree.fill_uniform_random()
                tree.compute_ev()
            d = tree.get_tree_as_dict()
            path = os.path.join(OUT, "export_%s_%s.json.gz" % (game, state))
            with gzip.GzipFile(path, "wb", mtime=0) as f:
                f.write(json.dumps(d).encode())
            print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
