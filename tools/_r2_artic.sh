cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_articulated.py -q -s 2>&1 | tail -25
python - <<'PY'
import sys, os, time, json
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import torch
from keypoints2body_b200 import synthetic as syn
from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter
from keypoints2body_b200.core.fitters.misc_models import FLAMEFitter, MANOFitter
from keypoints2body_b200.models.smpl_data import FLAMEData, MANOData, SMPLHData, SMPLXData
from oracle_free_problem import articulated_problem
gmm = syn.make_gmm(0); dev = torch.device("cuda")
for warp in ("1", "0"):
    os.environ["K2B_ARTIC_WARP"] = warp
    for mt in ("smplx", "smplh", "mano", "flame"):
        wmod, tgt_a, idx_a, init_a = articulated_problem(mt, 1, seed=700)
        cls = {"smplx": SMPLXData, "smplh": SMPLHData, "mano": MANOData, "flame": FLAMEData}[mt]
        for lb in (True, False):
            if mt in ("smplx", "smplh"):
                fit = WorldSpaceFitter(wmod, joints_category="GENERIC", model_type=mt, gmm=gmm, use_lbfgs=lb, device=dev)
            else:
                fit = (MANOFitter if mt == "mano" else FLAMEFitter)(wmod, coordinate_mode="world", use_lbfgs=lb, device=dev)
            p0 = cls(**{k: v.to(dev) for k, v in init_a.items()})
            for _ in range(2): r = fit.fit_frame(p0, tgt_a, torch.ones(len(idx_a)), seq_ind=0, target_model_indices=idx_a)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5): r = fit.fit_frame(p0, tgt_a, torch.ones(len(idx_a)), seq_ind=0, target_model_indices=idx_a)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
            print(f"warp={warp} {mt} {'lbfgs' if lb else 'adam'}: {dt*1e3:.2f} ms per fit_frame, loss {float(r.loss):.1f}", flush=True)
PY
