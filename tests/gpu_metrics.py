"""Side channel of the GPU tests: every test that measures something (a pixel fraction, an RMSE ratio, a ulp distance) also
appends it to gpurun_out/test_metrics.jsonl, so that the numbers behind the assertions can be committed under profiles/."""
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(test, **values):
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        clean = {}
        for k, v in values.items():
            try:
                clean[k] = float(v) if not isinstance(v, (str, bool, int, list, dict)) else v
            except Exception:        # noqa: BLE001
                clean[k] = str(v)
        with open(os.path.join(ROOT, "gpurun_out", "test_metrics.jsonl"), "a") as f:
            f.write(json.dumps({"test": test, "t": time.time(), **clean}) + "\n")
    except Exception:                # noqa: BLE001
        pass
    return values
