"""BASELINE configs[4]: throughput sweep over proposals per unit (100..1000) x queries per image (1..16), on N GPUs, next to the
reference's CPU path.

    python tools/sweep.py [fp32|tf32|bf16] [images_per_gpu]                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py fp32 2   (8 GPUs)

Every (image, query) is one unit with its own C4 map and RPN outputs (co-attention runs before the RPN in the reference,
faster_rcnn_coatt_transformer_sk.py:234-247).  Each rank owns `images_per_gpu` images x Q queries (no data-path collective);
a point is timed with CUDA events on every rank (3 warm-up + 5 timed passes, device-resident inputs) and reported as the
whole-job pairs/s over the MAX rank time.  The CPU column is the reference's CPU path (the reference's own C++ nms / roi_align
where oracle/_ref is built + the torch-CPU port of the head, all host cores) on ONE unit of P proposals, timed on rank 0 --
its per-pair throughput does not depend on Q.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ait_b200 import synth  # noqa: E402
from ait_b200.proposal import propose_rois  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
images = int(sys.argv[2]) if len(sys.argv) > 2 else 2
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
PS, QS = (100, 200, 300, 500, 1000), (1, 2, 4, 8, 16)

head = synth.spread_score_layer(synth.make_head(seed=0, calibrated=True, randomize_bn=True, compute_dtype=mode)).to(dev)
eng = head.engine()
max_units = images * 16
first = rank * max_units
maps = torch.stack([synth.c4_map(first + u) for u in range(max_units)]).to(dev)
qrys = torch.stack([synth.query_feat(first + u) for u in range(max_units)]).to(dev)
rpn = [synth.rpn_outputs(first + u) for u in range(max_units)]
boxes = torch.stack([r[0] for r in rpn]).to(dev)
scores = torch.stack([r[1] for r in rpn]).to(dev)

cpu = {}
if rank == 0 and "--no-cpu" not in sys.argv:
    sys.path.insert(0, ROOT)
    import bench  # noqa: E402  (the reference arm's step factory)
    for P in PS:
        step, kind, what = bench.cpu_step_factory(1, P)
        step()
        t0 = time.perf_counter()
        step()
        cpu[P] = P / (time.perf_counter() - t0)
    cpu_what = what

rows = []
for P in PS:
    for Q in QS:
        U = images * Q

        def step():
            rois, _ = propose_rois(boxes[:U], scores[:U], 6000, P, 0.7)
            return eng.head_forward(maps[:U], qrys[:U], rois)

        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(5):
            cls, _ = step()
        en.record()
        torch.cuda.synchronize()
        ms = st.elapsed_time(en) / 5
        ok = bool(torch.isfinite(cls).all())
        if world > 1:
            t = torch.tensor([ms, 0.0 if ok else 1.0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ok = float(t[0]), float(t[1]) == 0.0
        if rank == 0:
            row = dict(mode=mode, n_gpus=world, images_per_gpu=images, queries=Q, proposals=P, units=U * world,
                       pairs=U * P * world, ms=round(ms, 3), pairs_per_s=round(world * U * P / (ms * 1e-3)), finite=ok)
            if P in cpu:
                row["cpu_pairs_per_s"] = round(cpu[P], 1)
            rows.append(row)
            print(json.dumps(row), flush=True)
if rank == 0:
    print("# BASELINE configs[4] sweep (tools/sweep.py), %s configuration, %d x B200, %d images x Q queries x P proposals per GPU"
          % (mode, world, images))
    print("# cell: whole-job pairs/s (ms per batch, max over ranks); device-resident inputs, CUDA events, 3 warm-up + 5 timed passes")
    if cpu:
        print("# cpu: the reference's CPU path on this box's %d host cores, one unit of P proposals (%s)" % (os.cpu_count(), cpu_what))
    print("%-6s" % "P\\Q" + "".join("%20d" % q for q in QS) + ("%14s" % "cpu pairs/s" if cpu else ""))
    for P in PS:
        cells = [r for r in rows if r["proposals"] == P]
        print("%-6d" % P + "".join("%11d (%6.2f)" % (c["pairs_per_s"], c["ms"]) for c in cells) + ("%14.1f" % cpu[P] if cpu else ""))
if world > 1:
    dist.destroy_process_group()
