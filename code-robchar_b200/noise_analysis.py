"""Path grammar of the experiment files (upstream noise_analysis.py:33-61).

Only the pieces the MC path needs: ``ExperimentNamer`` (so cached ``.le/.mc/.mcm`` files resolve
to the same names) and the two exception types.  The experiment *driver* (optimiser runs) is out of
scope (orchestration).
"""
import os
from dataclasses import dataclass


@dataclass
class ExperimentNamer:
    experiment_name: str = "alpha"
    Nspin: int = 5
    inspin: int = 0
    outspin: int = 2
    numcontrollers: int = 100
    global_dir: str = "experiments"

    def home(self):
        home = self.global_dir + "/" + self.experiment_name
        if not os.path.exists(home):
            os.makedirs(home, exist_ok=True)
        return home

    def __call__(self):
        return f"{self.home()}/ppo_spin_{self.Nspin}_{self.inspin}-{self.outspin}_c_{self.numcontrollers}"


class ModelDoesNotExistError(Exception):
    def __init__(self):
        self.message = "Model not found in the current database!"
        super().__init__(self.message)


class DirectoryDoesNotExistError(Exception):
    def __init__(self, global_exp_path):
        self.message = "Directory not found in {}!".format(global_exp_path)
        super().__init__(self.message)
