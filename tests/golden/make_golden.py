#!/usr/bin/env python
"""Build the committed fixtures from the reference's own test data.

Run in the development container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Outputs (committed; they travel to the GPU box, /root/reference does not):

* tests/golden/toy.npz        -- the reference's 3-tank test problem
      inputs : src/test/testDataFiles/{network,scenarioTree,controllerConfig,forecastor}.json
      golden : src/test/testDataFiles/engineTest.json  (keys prefixed "engine.")
               src/test/testDataFiles/smpcTest.json    (keys prefixed "smpc.")
  These are the vectors the reference's own tests assert against
  (src/test/Testing.cu:340-531, src/test/TestSmpcController.cu:114-398).
* data/barcelona_base.npz     -- the Barcelona-size pieces that ship with the reference
      src/paser/dataSource/controllerConfig32.json (L, Lhat, W, preconditioner, state, prices)
      src/paser/dataSource/scenarioTree32.json, scenarioTree65.json (real trees, K=6 / K=30)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from rapidnet_b200.problem import (Config, Forecast, Network, Problem, Tree,  # noqa: E402
                                   problem_to_npz_dict)

REF = os.environ.get("RAPIDNET_REFERENCE", "/root/reference")
TD = os.path.join(REF, "src/test/testDataFiles")
DS = os.path.join(REF, "src/paser/dataSource")


def jl(path):
    with open(path) as fh:
        return json.load(fh)


def main():
    prob = Problem(
        network=Network.from_doc(jl(os.path.join(TD, "network.json"))),
        tree=Tree.from_doc(jl(os.path.join(TD, "scenarioTree.json"))),
        config=Config.from_doc(jl(os.path.join(TD, "controllerConfig.json"))),
        forecast=Forecast.from_doc(jl(os.path.join(TD, "forecastor.json"))),
    )
    out = problem_to_npz_dict(prob)
    for prefix, name in (("engine", "engineTest.json"), ("smpc", "smpcTest.json")):
        for k, v in jl(os.path.join(TD, name)).items():
            out[f"{prefix}.{k}"] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "toy.npz"), **out)
    print("wrote", os.path.join(HERE, "toy.npz"), len(out), "arrays")

    cfg = jl(os.path.join(DS, "controllerConfig32.json"))
    base = {}
    for k, v in cfg.items():
        if isinstance(v, list):
            base[f"cfg.{k}"] = np.asarray(v, dtype=np.float32 if len(v) > 1 else np.float64)
    for tag in ("32", "65"):
        t = Tree.from_doc(jl(os.path.join(DS, f"scenarioTree{tag}.json")))
        for k, v in t.__dict__.items():
            base[f"tree{tag}.{k}"] = np.asarray(v)
    os.makedirs(os.path.join(ROOT, "data"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "data", "barcelona_base.npz"), **base)
    print("wrote data/barcelona_base.npz", len(base), "arrays")


if __name__ == "__main__":
    main()
