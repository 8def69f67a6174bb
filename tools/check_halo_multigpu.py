"""torchrun check: schedule S2 sharded over N ranks (NCCL halo exchange at shard boundaries) must equal
the same sequence fitted on one rank.  Launch: python -m torch.distributed.run --nproc-per-node N ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from keypoints2body_b200 import synthetic as syn
from keypoints2body_b200.api.batch import SequenceBatchFitter
from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter
from keypoints2body_b200.distributed import shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T = 1000                                     # one sequence, boundaries fall mid-sequence
w = syn.make_body_model("smpl", seed=0)
mo = syn.make_motion(T, seed=9)
tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
seq = torch.arange(T, dtype=torch.int32)
ok = True
for lbfgs in (False, True):
    f = WorldSpaceFitter(w, joints_category="AMASS", use_lbfgs=lbfgs, model_type="smpl", gmm=syn.make_gmm(0), device=dev)
    lo, hi = shard_range(T, rank, world)
    mine = SequenceBatchFitter(f, hi - lo, with_vertices=False).run(tgt[lo:hi], seq[lo:hi])
    pose = mine["pose"].clone()
    gathered = [torch.empty(shard_range(T, r, world)[1] - shard_range(T, r, world)[0], 72, device=dev) for r in range(world)]
    dist.all_gather(gathered, pose)
    if rank == 0:
        whole = SequenceBatchFitter(f, T, with_vertices=False)
        # single rank: no process-group exchange (group of one behaves like no halo)
        ref = whole.run(tgt, seq, group=dist.new_group([0]) if False else None)["pose"] if world == 1 else None
    # the single-rank reference must not exchange: run it outside the group on every rank, compare on rank 0
    dist.barrier()
    os.environ["K2B_NO_HALO"] = "1"
    import keypoints2body_b200.api.batch as B
    orig = B.exchange_halo
    B.exchange_halo = lambda last, group=None: None
    ref = SequenceBatchFitter(f, T, with_vertices=False).run(tgt, seq)["pose"]
    B.exchange_halo = orig
    got = torch.cat(gathered)
    d = (got - ref).abs().max().item()
    same = torch.equal(got, ref)
    if rank == 0:
        print(f"{'lbfgs' if lbfgs else 'adam'}: sharded over {world} ranks vs single rank: max|dpose| = {d:.3e}, bit-identical = {same}")
    ok = ok and same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
