"""The worker scripts behind tests/test_gpu_{bonded,settle,pme,langevin}.py, run here against an oracle-backed stand-in
engine (tests/mock_engine.py): their Python -- workload keys, array shapes, comparisons, thresholds -- must already be
sound when they first meet a GPU, so that a failure there says something about the kernels."""
import importlib
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name,must_pass", [("bonded_gpu_worker", True), ("settle_gpu_worker", True), ("pme_gpu_worker", True),
                                            ("langevin_gpu_worker", False)])
def test_worker_script_runs_against_the_stand_in(name, must_pass, capsys, oracle):
    sys.path.insert(0, HERE)
    from mock_engine import MockEngine
    mod = importlib.import_module(name)
    mod.MdEngine = MockEngine
    rc = mod.main()
    res = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    if must_pass:
        assert rc == 0, res
    else:
        # the stand-in restarts the thermostat's step counter at every step() call, so the long thermalisation leg repeats
        # its noise; the same-noise trajectory legs are exact
        assert res["traj_ok"] and res["csvr_traj_ok"] and res["dv"] < 5e-3 and res["csvr_dv"] < 5e-3, res
