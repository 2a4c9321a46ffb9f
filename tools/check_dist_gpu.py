"""Multi-GPU parity check (run under torchrun on N GPUs): the destination-partitioned CUDA path must match
the single-GPU CUDA path on the same graph -- bit-equal top-k ids, floats within BASELINE.json's tolerances
(1e-5 fp32, 1e-2 bf16; no slack factors) -- for both sharded paths (tape-free overlapped, autograd tape),
both dtypes, and for a counter-based graph sharded without materialising it.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
      --master-port 29511 tools/check_dist_gpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from truth_recommendation_gnn_b200 import dist_check  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    bad = []
    for dtype in (torch.float32, torch.bfloat16):
        for fused, gen in ((True, False), (False, False), (True, True)):
            r = dist_check.check_sharded_against_single(dev, dtype, fused=fused, from_generator=gen)
            if rank == 0:
                print(("dist parity ok: " if r["ok"] else "dist parity FAILED: ") +
                      " ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in r.items()), flush=True)
            if not r["ok"]:
                bad.append(r)
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
