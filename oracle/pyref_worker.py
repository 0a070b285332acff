"""TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/README.md): times the UNMODIFIED Python reference's MapEnv.step in one
process -- BASELINE.md section 4 / SURVEY.md 8d config 1: `HarvestEnv(num_agents=5)` (or CleanupEnv), default map,
np.random.seed(s); random.seed(s); reset(); `steps` steps of actions np.random.RandomState(s).randint(A, size=(steps, N))
passed as {'agent-j': int} in index order; only env.step is timed.  bench.py starts one of these per host core.

    python oracle/pyref_worker.py --root <reference tree> --game harvest --agents 5 --view 7 --seed 0 --steps 1000

Prints one JSON line {"env_steps_per_s": ..., "agent_steps_per_s": ..., "steps": ..., "seconds": ...}."""
import argparse
import json
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", required=True)
    ap.add_argument("--game", default="harvest")
    ap.add_argument("--agents", type=int, default=5)
    ap.add_argument("--view", type=int, default=7)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1000)
    args = ap.parse_args()
    os.environ["SSD_REFERENCE_ROOT"] = args.root
    import numpy as np
    from oracle import ref_harness as rh
    rh.REFERENCE_ROOT = args.root
    ref = rh.load_reference()
    ref.harvest.HARVEST_VIEW_SIZE = args.view     # read as module globals when the env is built (harvest.py:54, cleanup.py:127)
    ref.cleanup.CLEANUP_VIEW_SIZE = args.view
    np.random.seed(args.seed)
    random.seed(args.seed)
    cls = ref.HarvestEnv if args.game == "harvest" else ref.CleanupEnv
    env = cls(num_agents=args.agents)
    env.reset()
    n_act = 8 if args.game == "harvest" else 9
    acts = np.random.RandomState(args.seed).randint(n_act, size=(args.steps, args.agents))
    ids = ['agent-%d' % j for j in range(args.agents)]
    dicts = [{ids[j]: int(a[j]) for j in range(args.agents)} for a in acts]
    t0 = time.perf_counter()
    for d in dicts:
        env.step(d)
    dt = time.perf_counter() - t0
    print(json.dumps({"env_steps_per_s": args.steps / dt, "agent_steps_per_s": args.steps * args.agents / dt,
                      "steps": args.steps, "seconds": dt}))


if __name__ == "__main__":
    main()
