"""Prints the per-fixture latent-state error margins of the CUDA path against the reference-run golden fixtures."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sf_oracle as so
from tests.test_gpu_rollout import _nnfo, _rel, TOL

for precision in ("bf16", "bf16x3"):
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "c64_latent_*.npz"))):
        z = np.load(path)
        C, H, seed = int(z["C"]), int(z["H"]), int(z["seed"])
        m = _nnfo(str(z["solver"]), bool(z["variable"]), bool(z["impute"]), seed, float(z["gain"]), precision)
        m.record_all = True
        times, targets = z["times"].tolist(), z["targets"].tolist()
        obs = so.recipe_array("obs", (1, len(times), C, H, H), seed).cuda()
        tape = torch.stack([so.recipe_array(f"eps{i}", (C, H // 4, H // 4), seed) for i in range(int(z["n_eps"]))]).cuda()
        m._draw_noise = lambda n, h, w, device: tape[:max(n, 1)].contiguous()
        with torch.no_grad():
            state, aux, x = m(torch.tensor(times, dtype=torch.float64), torch.zeros(1, 1, C, H, H, device="cuda"), obs, 0.05,
                              torch.tensor(targets, dtype=torch.float64))
        ref = torch.from_numpy(z["states_f64"])
        got = m.last_trace[0].cpu()
        errs = [_rel(got[i], ref[i]) for i in range(ref.shape[0])]
        print(f"{precision:7s} {os.path.basename(path)[11:-4]:28s} max per-event latent err {max(errs):.3e} (tol {TOL[precision]:.0e})  "
              f"decoded {_rel(x[:, [0, -1]].cpu(), torch.from_numpy(z['x_f64'])):.3e} (tol {5 * TOL[precision]:.0e})")
