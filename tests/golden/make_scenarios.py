"""Runs tests/scenarios.py against the UNMODIFIED Python reference at /root/reference and stores every
record in tests/golden/scenarios.npz (build container only; see oracle/ref_harness.py for the import
stubs).  tests/test_adapter_gpu.py replays the same scenarios through the adapters on the GPU."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as rh  # noqa: E402
import scenarios  # noqa: E402


def reference_api():
    ref = rh.load_reference()
    return types.SimpleNamespace(MapEnv=ref.MapEnv, HarvestEnv=ref.HarvestEnv, CleanupEnv=ref.CleanupEnv,
                                 Agent=ref.agent.Agent, HarvestAgent=ref.agent.HarvestAgent,
                                 CleanupAgent=ref.agent.CleanupAgent, BASE_ACTIONS=ref.agent.BASE_ACTIONS,
                                 HARVEST_ACTIONS=ref.agent.HARVEST_ACTIONS, CLEANUP_ACTIONS=ref.agent.CLEANUP_ACTIONS)


def main():
    res = scenarios.run_all(reference_api())
    flat = {}
    for sc, items in res.items():
        names = []
        for i, (name, v) in enumerate(items):
            flat["%s|%05d" % (sc, i)] = v
            names.append(name)
        flat[sc + "|names"] = np.array(names)
        print("%-24s %5d records" % (sc, len(items)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenarios.npz")
    np.savez_compressed(path, **flat)
    print("%s  %.1f KB" % (path, os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
