"""Problem containers and JSON I/O in the reference's schema.

The four input documents of the reference are plain JSON objects whose scalars
are 1-element arrays and whose matrices are flat column-major arrays:

* network        -- /root/reference/src/DwnNetwork.cuh:23-37, DwnNetwork.cu:43-112
* scenario tree  -- /root/reference/src/ScenarioTree.cuh:23-40, ScenarioTree.cu:46-121
* forecaster     -- /root/reference/src/Forecaster.cuh:23-30, Forecaster.cu:38-119
                    (positional: member 4+2t is the demand of slot t, 5+2t its prices)
* controller cfg -- /root/reference/src/SmpcConfiguration.cuh:24-47, SmpcConfiguration.cu:44-119

This module is host-side plumbing (tests, generators, bench); the product's own
loaders are the C++ classes in rapidnet_b200/host/.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

F32 = np.float32
I32 = np.int32


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=F32).reshape(-1))


def _i(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1).astype(I32))


def _scalar(doc, key) -> float:
    v = doc[key]
    return float(v[0]) if isinstance(v, list) else float(v)


def _num_list(a) -> list:
    a = np.asarray(a)
    if a.dtype.kind in "iu":
        return [int(x) for x in a.reshape(-1)]
    # repr of the float32 value (shortest string that round-trips the float32)
    return [float(np.format_float_positional(x, unique=True, trim="0")) if np.isfinite(x) else float(x)
            for x in a.astype(F32).reshape(-1)]


@dataclass
class Network:
    nx: int
    nu: int
    nd: int
    ne: int
    A: np.ndarray      # nx*nx   (loaded, never used by the reference: A = I)
    B: np.ndarray      # nx*nu   col-major
    Gd: np.ndarray     # nx*nd
    E: np.ndarray      # ne*nu
    Ed: np.ndarray     # ne*nd
    xmin: np.ndarray
    xmax: np.ndarray
    xsafe: np.ndarray
    umin: np.ndarray
    umax: np.ndarray
    alpha1: np.ndarray
    N: int = 24

    @staticmethod
    def from_doc(d) -> "Network":
        return Network(
            nx=int(_scalar(d, "nx")), nu=int(_scalar(d, "nu")), nd=int(_scalar(d, "nd")),
            ne=int(_scalar(d, "ne")), A=_f(d["matA"]), B=_f(d["matB"]), Gd=_f(d["matGd"]),
            E=_f(d["matE"]), Ed=_f(d["matEd"]), xmin=_f(d["vecXmin"]), xmax=_f(d["vecXmax"]),
            xsafe=_f(d["vecXsafe"]), umin=_f(d["vecUmin"]), umax=_f(d["vecUmax"]),
            alpha1=_f(d["costAlpha1"]), N=int(_scalar(d, "N")) if "N" in d else 24)

    def to_doc(self) -> dict:
        return {
            "nx": [self.nx], "nu": [self.nu], "ne": [self.ne], "nd": [self.nd], "N": [self.N],
            "matA": _num_list(self.A), "matB": _num_list(self.B), "matGd": _num_list(self.Gd),
            "matE": _num_list(self.E), "matEd": _num_list(self.Ed),
            "vecXmin": _num_list(self.xmin), "vecXmax": _num_list(self.xmax),
            "vecXsafe": _num_list(self.xsafe), "vecUmin": _num_list(self.umin),
            "vecUmax": _num_list(self.umax), "costAlpha1": _num_list(self.alpha1),
        }


@dataclass
class Tree:
    N: int
    K: int
    nodes: int
    n_nonleaf: int
    n_children_tot: int
    stages: np.ndarray                 # [nodes] 0-based stage
    nodes_per_stage: np.ndarray        # [N+1] trailing 0
    nodes_per_stage_cumul: np.ndarray  # [N+2]
    leaves: np.ndarray                 # [K] 1-based
    children: np.ndarray               # [n_children_tot] 1-based
    ancestor: np.ndarray               # [nodes] 1-based, root = 0
    n_children: np.ndarray             # [n_nonleaf]
    n_children_cumul: np.ndarray       # [nodes]
    prob: np.ndarray                   # [nodes]
    dim_demand: int
    dim_price: int
    err_demand: np.ndarray             # [nodes*nd] node-major
    err_price: np.ndarray              # [nodes*nu]

    @staticmethod
    def from_doc(d) -> "Tree":
        return Tree(
            N=int(_scalar(d, "N")), K=int(_scalar(d, "K")), nodes=int(_scalar(d, "nodes")),
            n_nonleaf=int(_scalar(d, "nNonLeafNodes")), n_children_tot=int(_scalar(d, "nChildrenTot")),
            stages=_i(d["stages"]), nodes_per_stage=_i(d["nodesPerStage"]),
            nodes_per_stage_cumul=_i(d["nodesPerStageCumul"]), leaves=_i(d["leaves"]),
            children=_i(d["children"]), ancestor=_i(d["ancestor"]), n_children=_i(d["nChildren"]),
            n_children_cumul=_i(d["nChildrenCumul"]), prob=_f(d["probNode"]),
            dim_demand=int(_scalar(d, "dimDemand")), dim_price=int(_scalar(d, "dimPrice")),
            err_demand=_f(d["errorDemandNode"]), err_price=_f(d["errorPriceNode"]))

    def to_doc(self) -> dict:
        return {
            "N": [self.N], "K": [self.K], "dimDemand": [self.dim_demand], "dimPrice": [self.dim_price],
            "nodes": [self.nodes], "nChildrenTot": [self.n_children_tot],
            "nNonLeafNodes": [self.n_nonleaf], "stages": _num_list(self.stages),
            "nodesPerStage": _num_list(self.nodes_per_stage),
            "nodesPerStageCumul": _num_list(self.nodes_per_stage_cumul),
            "leaves": _num_list(self.leaves), "children": _num_list(self.children),
            "ancestor": _num_list(self.ancestor), "nChildren": _num_list(self.n_children),
            "nChildrenCumul": _num_list(self.n_children_cumul), "probNode": _num_list(self.prob),
            "errorDemandNode": _num_list(self.err_demand), "errorPriceNode": _num_list(self.err_price),
        }

    # /root/reference/src/ScenarioTree.cu:149-169
    def final_branch_node(self) -> int:
        nps, cum = self.nodes_per_stage, self.nodes_per_stage_cumul
        for s in range(self.N - 1):
            if nps[s] == nps[s + 1]:
                return int(cum[s + 1])
        return 0

    def final_branch_stage(self) -> int:
        nps = self.nodes_per_stage
        for s in range(self.N - 1):
            if nps[s] == nps[s + 1]:
                return s
        return 0


@dataclass
class Config:
    nx: int
    nu: int
    nd: int
    nv: int
    N: int
    L: np.ndarray          # nu*nv (parsed by the reference, recomputed by its Engine)
    Lhat: np.ndarray       # nu*nd
    costW: np.ndarray      # nu*nu
    penalty_x: float
    penalty_xs: float
    precond: np.ndarray    # N*(nu+2nx), per stage [u | x | xsafe]
    current_x: np.ndarray
    prev_u: np.ndarray
    prev_demand: np.ndarray
    step_size: float
    max_iter: int
    path_network: str = ""
    path_tree: str = ""
    path_forecaster: str = ""
    algorithm: str = "proximalAlgorithm"
    lbfgs: int = 5
    ne: Optional[int] = None

    @staticmethod
    def from_doc(d) -> "Config":
        return Config(
            nx=int(_scalar(d, "nx")), nu=int(_scalar(d, "nu")), nd=int(_scalar(d, "nd")),
            nv=int(_scalar(d, "nv")), N=int(_scalar(d, "N")), L=_f(d["matL"]), Lhat=_f(d["matLhat"]),
            costW=_f(d["costW"]), penalty_x=_scalar(d, "penaltyStateX"),
            penalty_xs=_scalar(d, "penaltySafetyX"), precond=_f(d["matDiagPrecnd"]),
            current_x=_f(d["currentX"]), prev_u=_f(d["prevU"]), prev_demand=_f(d["prevDemand"]),
            step_size=_scalar(d, "stepSize"), max_iter=int(_scalar(d, "maxIterations")),
            path_network=d.get("pathToNetwork", ""), path_tree=d.get("pathToScenarioTree", ""),
            path_forecaster=d.get("pathToForecaster", ""),
            algorithm=d.get("algorithmName", "proximalAlgorithm"),
            lbfgs=int(_scalar(d, "lbfgsBufferSize")) if "lbfgsBufferSize" in d else 5,
            ne=int(_scalar(d, "ne")) if "ne" in d else None)

    def to_doc(self) -> dict:
        doc = {
            "nx": [self.nx], "nu": [self.nu], "nv": [self.nv], "nd": [self.nd], "N": [self.N],
            "matL": _num_list(self.L), "matLhat": _num_list(self.Lhat),
            "matDiagPrecnd": _num_list(self.precond), "costW": _num_list(self.costW),
            "currentX": _num_list(self.current_x), "prevDemand": _num_list(self.prev_demand),
            "prevU": _num_list(self.prev_u), "stepSize": [float(self.step_size)],
            "maxIterations": [int(self.max_iter)], "penaltyStateX": [float(self.penalty_x)],
            "penaltySafetyX": [float(self.penalty_xs)], "pathToNetwork": self.path_network,
            "pathToScenarioTree": self.path_tree, "pathToForecaster": self.path_forecaster,
            "algorithmName": self.algorithm, "lbfgsBufferSize": [int(self.lbfgs)],
        }
        if self.ne is not None:
            doc["ne"] = [self.ne]
        return doc


@dataclass
class Forecast:
    N: int
    sim_horizon: int
    dim_demand: int
    dim_prices: int
    demand: List[np.ndarray] = field(default_factory=list)   # per slot, N*nd (stage-major)
    prices: List[np.ndarray] = field(default_factory=list)   # per slot, N*nu
    names: List[str] = field(default_factory=list)

    @staticmethod
    def from_doc(d) -> "Forecast":
        keys = list(d.keys())
        fc = Forecast(N=int(_scalar(d, "N")), sim_horizon=int(_scalar(d, "simHorizon")),
                      dim_demand=int(_scalar(d, "dimDemand")), dim_prices=int(_scalar(d, "dimPrices")))
        t = 0
        # positional semantics: /root/reference/src/Forecaster.cu:93-119
        while 5 + 2 * t < len(keys):
            fc.demand.append(_f(d[keys[4 + 2 * t]]))
            fc.prices.append(_f(d[keys[5 + 2 * t]]))
            fc.names += [keys[4 + 2 * t], keys[5 + 2 * t]]
            t += 1
        return fc

    def to_doc(self) -> dict:
        doc = {"N": [self.N], "simHorizon": [self.sim_horizon], "dimDemand": [self.dim_demand],
               "dimPrices": [self.dim_prices]}
        for t, (dm, pr) in enumerate(zip(self.demand, self.prices)):
            doc[f"timeIdDemand{4875 + t}"] = _num_list(dm)
            doc[f"timeIdPrice{4875 + t}"] = _num_list(pr)
        return doc


@dataclass
class Problem:
    network: Network
    tree: Tree
    config: Config
    forecast: Forecast

    @property
    def dims(self) -> Dict[str, int]:
        n, t, c = self.network, self.tree, self.config
        return dict(nx=n.nx, nu=n.nu, nd=n.nd, ne=n.ne, nv=c.nv, N=t.N, K=t.K, nodes=t.nodes)


def load_json(path: str) -> dict:
    with open(path, "r") as fh:
        return json.load(fh)


def load_problem(config_path: str, base_dir: Optional[str] = None) -> Problem:
    """Load a controller config and the three documents it points to.

    Paths inside the config are resolved relative to `base_dir` (default: the
    process cwd, which is what the reference does -- Engine.cu:128-132)."""
    cfg = Config.from_doc(load_json(config_path))

    def _res(p):
        return p if os.path.isabs(p) or base_dir is None else os.path.join(base_dir, p)

    net = Network.from_doc(load_json(_res(cfg.path_network)))
    tree = Tree.from_doc(load_json(_res(cfg.path_tree)))
    fc = Forecast.from_doc(load_json(_res(cfg.path_forecaster)))
    return Problem(net, tree, cfg, fc)


def write_problem(p: Problem, out_dir: str, absolute_paths: bool = True) -> str:
    """Write the four JSON documents in the reference schema; returns the config path."""
    os.makedirs(out_dir, exist_ok=True)
    names = {"network": "network.json", "tree": "scenarioTree.json", "forecast": "forecastor.json"}
    root = os.path.abspath(out_dir) if absolute_paths else out_dir
    p.config.path_network = os.path.join(root, names["network"])
    p.config.path_tree = os.path.join(root, names["tree"])
    p.config.path_forecaster = os.path.join(root, names["forecast"])
    for obj, name in ((p.network, names["network"]), (p.tree, names["tree"]), (p.forecast, names["forecast"])):
        with open(os.path.join(out_dir, name), "w") as fh:
            json.dump(obj.to_doc(), fh)
    cfg_path = os.path.join(out_dir, "controllerConfig.json")
    with open(cfg_path, "w") as fh:
        json.dump(p.config.to_doc(), fh)
    return cfg_path


# ---- npz packing (fixtures travel to the GPU box as .npz, not as reference files) ----

def _pack(prefix: str, obj, out: dict):
    for k, v in obj.__dict__.items():
        if isinstance(v, list):
            if v and isinstance(v[0], np.ndarray):
                out[f"{prefix}.{k}"] = np.stack(v)
            else:
                out[f"{prefix}.{k}"] = np.array(v)
        elif v is None:
            continue
        else:
            out[f"{prefix}.{k}"] = np.asarray(v)


def problem_to_npz_dict(p: Problem) -> dict:
    out: dict = {}
    _pack("network", p.network, out)
    _pack("tree", p.tree, out)
    _pack("config", p.config, out)
    _pack("forecast", p.forecast, out)
    return out


def problem_from_npz_dict(z) -> Problem:
    def grab(prefix, cls):
        kw = {}
        for name, f in cls.__dataclass_fields__.items():
            key = f"{prefix}.{name}"
            if key not in z:
                continue
            v = z[key]
            t = f.type
            if t in ("int", int):
                kw[name] = int(v)
            elif t in ("float", float):
                kw[name] = float(v)
            elif t in ("str", str):
                kw[name] = str(v)
            elif t.startswith("List[np.ndarray]"):
                kw[name] = [np.ascontiguousarray(r) for r in v]
            elif t.startswith("List[str]"):
                kw[name] = [str(s) for s in v]
            elif t.startswith("Optional[int]"):
                kw[name] = int(v)
            else:
                kw[name] = np.ascontiguousarray(v)
        return cls(**kw)

    return Problem(grab("network", Network), grab("tree", Tree), grab("config", Config),
                   grab("forecast", Forecast))
