"""tests/golden/gen_golden_models.py -- golden forward passes of the reference's OWN model classes
(models/ac_models_hetero.py imported unmodified under the ray stubs of oracle/ref_harness.py).

Weights are filled from a numpy PCG64 stream (hhmarl_2d_b200.models.fill_from_seed) so the tests
can rebuild the same weights anywhere; the fixture stores only inputs and outputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402
from hhmarl_2d_b200.models import fill_from_seed  # noqa: E402


def main():
    m = rh.reference_models()
    out = {}
    rng = np.random.default_rng(7)
    for name, n_out in (("Fight1", 26), ("Fight2", 24), ("Esc1", 26), ("Esc2", 24)):
        torch.manual_seed(0)
        model = getattr(m, name)(None, None, n_out, {}, name)
        fill_from_seed(model, 100 + n_out + len(name))
        ac1 = name.endswith("1")
        esc = name.startswith("Esc")
        own = (30 if ac1 else 29) if esc else (26 if ac1 else 24)
        oth = (29 if ac1 else 30) if esc else (24 if ac1 else 26)
        a_own, a_oth = (4, 3) if ac1 else (3, 4)
        for tag, B, T in (("t1", 12, 1), ("t5", 15, 5)):
            obs = {"obs_1_own": rng.random((B, own), dtype=np.float32), "obs_2": rng.random((B, oth), dtype=np.float32),
                   "act_1_own": rng.random((B, a_own), dtype=np.float32), "act_2": rng.random((B, a_oth), dtype=np.float32)}
            with torch.no_grad():
                logits, _ = model({"obs": {k: torch.from_numpy(v) for k, v in obs.items()}}, [torch.tensor(0)],
                                  torch.tensor([T] * (B // T)))
                val = model.value_function()
            for k, v in obs.items():
                out[f"{name}_{tag}_{k}"] = v
            out[f"{name}_{tag}_logits"] = logits.numpy()
            out[f"{name}_{tag}_value"] = val.numpy()
    # CommanderGru (models/ac_models_hier.py), imported unmodified under the same stubs
    import models.ac_models_hier as mh
    torch.manual_seed(0)
    cm = mh.CommanderGru(None, None, 3, {}, "commander")
    fill_from_seed(cm, 777)
    for tag, B, T in (("t1", 8, 1), ("t4", 12, 4)):
        obs = {k: rng.random((B, 34), dtype=np.float32) for k in ("obs_1_own", "obs_2", "obs_3")}
        obs.update({k: rng.integers(0, 3, (B, 1)).astype(np.float32) for k in ("act_1_own", "act_2", "act_3")})
        nseq = B // T
        state = [torch.from_numpy(rng.random((nseq, 200), dtype=np.float32)) for _ in range(2)]
        with torch.no_grad():
            logits, new_state = cm({"obs": {k: torch.from_numpy(v) for k, v in obs.items()}}, state,
                                   torch.tensor([T] * nseq))
            val = cm.value_function()
        for k, v in obs.items():
            out[f"Cmd_{tag}_{k}"] = v
        out[f"Cmd_{tag}_h0"], out[f"Cmd_{tag}_h1"] = state[0].numpy(), state[1].numpy()
        out[f"Cmd_{tag}_logits"], out[f"Cmd_{tag}_value"] = logits.numpy(), val.numpy()
        out[f"Cmd_{tag}_nh0"], out[f"Cmd_{tag}_nh1"] = new_state[0].numpy(), new_state[1].numpy()
    path = os.path.join(ROOT, "tests", "golden", "models_forward.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
