"""torchrun worker: ONLY the data-parallel sub-record of bench.py (configs.dp4096) at the launched world size, one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 tools/dp_scale.py [--steps 24]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=24)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rec = bench.time_data_parallel(a.steps, 6, local, rank, world)
    if rank == 0:
        rec["n_gpus"] = world
        bench.emit(rec)        # bench redirects stdout to stderr on import; emit() writes the one JSON line to the real stdout
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
