"""N-GPU check of the sharded search path (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py
Validates, on every rank:
  (1) the shard's local GPU search == the CPU oracle searching the same (exported) shard graph, bit for bit;
  (2) the fused path (K1 epilogue pack -> ONE ncclAllGather -> merge kernel) == the host statement of the protocol
      (gloo all_gather of the local results + sharded.merge_keys), bit for bit, and is identical on all ranks.
Optionally times the sharded search (--bench)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "instant-distance_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from instant_distance_b200 import _abi, sharded  # noqa: E402
from tests import datagen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=200_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--queries", type=int, default=2000)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--ef", type=int, default=100)
ap.add_argument("--bench", type=int, default=0, help="timed repetitions of the sharded search")
ap.add_argument("--no-oracle", action="store_true")
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")


def exchange(uid):
    box = [uid]
    dist.broadcast_object_list(box, src=0)
    return box[0]


gen = datagen.sift_shaped
lo, hi = sharded.shard_range(args.points, rank, world)
rows = gen(args.points, args.dim, 5)[lo:hi]
q = gen(args.queries, args.dim, 6)
t0 = time.time()
sh = sharded.ShardedIndex(rows, lo, rank, world, local, exchange, seed=100 + rank)
build_s = time.time() - t0

ids, d, lens = sh.search(q, args.ef, args.k)

# (1) local search vs the oracle on this shard's graph
ok_local = True
sh.index.set_id_map(None)
l_ids, l_d, l_lens = sh.index.search(q, ef_search=args.ef, k=args.k)
if not args.no_oracle:
    from oracle import oracle as O

    p, zero, upper = sh.index.export_graph()
    ox = O.from_graph(O.Graph(p, zero, upper, 32, args.ef))
    o_ids, o_d, o_lens = ox.search(q, ef_search=args.ef, k=args.k, threads=8)
    ok_local = bool((o_ids == l_ids).all() and o_d.tobytes() == l_d.tobytes() and (o_lens == l_lens).all())

# (2) host statement of the protocol over gloo
# pid -> global id map, recomputed from the deterministic seeded shuffle (the GPU build's permutation equals the
# oracle's: tests/test_gpu_build.py::test_shuffle_matches_oracle)
from oracle import oracle as O2  # noqa: E402

local_ids = O2.shuffle(hi - lo, 100 + rank)
gmap = sharded.global_id_map(local_ids, lo)
gids = np.where(l_ids == 0xFFFFFFFF, 0, gmap[np.minimum(l_ids, hi - lo - 1)])
keys = sharded.pack_keys(l_d, gids, np.minimum(l_lens, args.k))
gathered = [torch.empty((args.queries, args.k), dtype=torch.int64) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(keys.view(np.int64)))
m_ids, m_d, m_lens = sharded.merge_keys(np.stack([g.numpy().view(np.uint64) for g in gathered]), args.k)
ok_merge = bool((m_ids == ids).all() and m_d.tobytes() == d.tobytes() and (m_lens == lens).all())
diag = {}
if not ok_merge:
    bad = np.nonzero((m_ids != ids).any(axis=1) | (m_lens != lens))[0]
    diag = {"rank": rank, "bad_rows": int(len(bad)), "dist_bits_equal": bool(m_d.tobytes() == d.tobytes()),
            "lens_equal": bool((m_lens == lens).all())}
    if len(bad):
        b = int(bad[0])
        diag.update(first=b, fused_ids=ids[b].tolist(), proto_ids=m_ids[b].tolist(), fused_d=d[b].tolist(), proto_d=m_d[b].tolist(),
                    fused_len=int(lens[b]), proto_len=int(m_lens[b]))
    print("DIAG", json.dumps(diag), file=sys.stderr, flush=True)

flags = torch.tensor([int(ok_local), int(ok_merge)])
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
res = {"world": world, "n": args.points, "shard": [lo, hi], "build_s": round(build_s, 2), "local_eq_oracle": bool(flags[0]),
       "fused_eq_protocol": bool(flags[1])}

if args.bench:
    sh.index.set_id_map(gmap)
    dq = torch.from_numpy(q).cuda()
    d_ids = torch.empty((args.queries, args.k), dtype=torch.int32, device="cuda")
    d_d = torch.empty((args.queries, args.k), dtype=torch.float32, device="cuda")
    d_l = torch.empty((args.queries,), dtype=torch.int32, device="cuda")
    stream = torch.cuda.ExternalStream(sh.index.stream, device=local)
    for _ in range(3):
        sh.index.sharded_search_device(sh.comm, dq.data_ptr(), args.queries, args.ef, args.k, d_ids.data_ptr(), d_d.data_ptr(), d_l.data_ptr())
    sh.index.sync()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.bench):
        sh.index.sharded_search_device(sh.comm, dq.data_ptr(), args.queries, args.ef, args.k, d_ids.data_ptr(), d_d.data_ptr(), d_l.data_ptr())
    e1.record(stream)
    e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.bench])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res["sharded_ms_per_batch"] = float(ms)
    res["sharded_qps"] = args.queries / (float(ms) / 1e3)
if rank == 0:
    print(json.dumps(res), flush=True)
sh.close()
dist.destroy_process_group()
sys.exit(0 if (flags[0] and flags[1]) else 1)
