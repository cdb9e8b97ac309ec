"""Run under torchrun (N >= 2): the range-sharded multi-GPU path must reproduce the
single-GPU factors.  Rank 0 prints the comparison."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import myrrix_recommender_b200 as M
from myrrix_recommender_b200.factorizer import comm_unique_id

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
import traceback
for (U, I, nnz, k) in ((3000, 500, 10, 16), (20001, 3001, 30, 32), (50000, 7000, 64, 64)):
  try:
      als = M.NativeALS(k, device=lr)
      uid = [comm_unique_id() if rank == 0 else None]
      dist.broadcast_object_list(uid, src=0)
      als.comm_init(rank, world, uid[0])
      als.synth_interactions(U, I, nnz, seed=1234567890, neg_fraction=0.05)
      als.synth_y0(seed=1234567890)
      als.iterate(3); als.sync()
      X, Y = als.get_x(), als.get_y()
      tm = als.timings()
      als.close()
      if rank == 0:
          ref = M.NativeALS(k, device=lr)
          ref.synth_interactions(U, I, nnz, seed=1234567890, neg_fraction=0.05)
          ref.synth_y0(seed=1234567890)
          ref.iterate(3); ref.sync()
          Xr, Yr = ref.get_x(), ref.get_y()
          ref.close()
          dx, dy = np.abs(X - Xr).max(), np.abs(Y - Yr).max()
          print("U=%d I=%d k=%d world=%d: max|X-Xref|=%.3g max|Y-Yref|=%.3g finite=%s" % (
              U, I, k, world, dx, dy, np.isfinite(X).all() and np.isfinite(Y).all()))
          ok = ok and dx <= 1e-5 * np.abs(Xr).max() and dy <= 1e-5 * np.abs(Yr).max()
      dist.barrier()
  except Exception:
    traceback.print_exc(); sys.stdout.flush(); os._exit(1)
# Host-buffer path: every rank uploads only its by-user block; the by-item blocks are built on the
# devices (all-to-all over NCCL); result vs the single-GPU build of the same matrix.
from oracle import synth
from myrrix_recommender_b200.sharding import local_block
for (U, I, nnz, k) in ((20001, 3001, 30, 64),):
  try:
      ptr, idx, val = synth.synth_rows(0, U, I, nnz, seed=7, neg_fraction=0.05)
      Y0 = synth.unit_rows(I, k, seed=7)
      ub, ue = local_block(U, rank, world)
      e0, e1 = int(ptr[ub]), int(ptr[ue])
      als = M.NativeALS(k, device=lr)
      uid = [comm_unique_id() if rank == 0 else None]
      dist.broadcast_object_list(uid, src=0)
      als.comm_init(rank, world, uid[0])
      als.set_interactions(U, I, ptr[ub:ue + 1] - e0, idx[e0:e1], val[e0:e1])
      als.set_y(Y0)
      als.iterate(3); als.sync()
      X, Y = als.get_x(), als.get_y()
      p2p = os.environ.get("MYRRIX_ALS_NO_P2P", "0")
      als.close()
      if rank == 0:
          ref = M.NativeALS(k, device=lr)
          ref.set_interactions(U, I, ptr, idx, val)
          ref.set_y(Y0)
          ref.iterate(3); ref.sync()
          Xr, Yr = ref.get_x(), ref.get_y()
          ref.close()
          dx, dy = np.abs(X - Xr).max(), np.abs(Y - Yr).max()
          print("host blocks + distributed transpose (NO_P2P=%s) U=%d I=%d k=%d world=%d: max|X-Xref|=%.3g max|Y-Yref|=%.3g"
                % (p2p, U, I, k, world, dx, dy))
          ok = ok and dx <= 1e-5 * np.abs(Xr).max() and dy <= 1e-5 * np.abs(Yr).max()
      dist.barrier()
  except Exception:
    traceback.print_exc(); sys.stdout.flush(); os._exit(1)
if rank == 0:
    print("MULTI-GPU CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
