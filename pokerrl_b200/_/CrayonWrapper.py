"""Log exporter with the call surface of `PokerRL/_/CrayonWrapper.py:12-93`.

The reference pushes ChiefBase's log buffer to a PyCrayon/TensorBoard docker server over HTTP; that external
service is out of scope.  This class keeps `update_from_log_buffer()` / `export_all(iter_nr)` so
that `examples/run_cfrp_example.py:22-37` runs unchanged, and writes the same information as JSON files when
`path_log_storage` is given."""
import json
import os


class CrayonWrapper:
    def __init__(self, name, runs_distributed, runs_cluster, chief_handle, path_log_storage=None,
                 crayon_server_address="localhost"):
        self._name = name
        self._chief_handle = chief_handle
        self._path_log_storage = path_log_storage
        if path_log_storage is not None:
            os.makedirs(path_log_storage, exist_ok=True)
        self._experiments = {}
        self.clear()

    @property
    def name(self):
        return self._name

    @property
    def path_log_storage(self):
        return self._path_log_storage

    def clear(self):
        self._experiments = {}
        self._custom_logs = {}

    def update_from_log_buffer(self):
        new_v, exp_names = self._chief_handle.get_new_values()
        for e in exp_names:
            self._custom_logs.setdefault(e, {})
        for exp, graphs in new_v.items():
            for graph, points in graphs.items():
                for step, value in points:
                    self._custom_logs[exp].setdefault(graph, []).append({step: value})

    def export_all(self, iter_nr):
        if self._path_log_storage is None:
            return
        d = os.path.join(self._path_log_storage, str(self._name), str(iter_nr), "as_json")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "logs.json"), "w") as f:
            json.dump(self._custom_logs, f)
