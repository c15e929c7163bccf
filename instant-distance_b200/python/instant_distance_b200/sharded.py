"""Host-side logic of the PointId-range-sharded index: which rows a rank owns, the per-shard build, the
global-id map, and the protocol around the single all-gather.  All numerics run in the C library; the functions here
that do not touch the GPU (`shard_range`, `pack_keys`, `merge_keys`) define the protocol and are what the world_size-2
gloo tests exercise on CPU."""
import numpy as np

from . import _abi

KEY_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_range(n, rank, world):
    """Contiguous range of the INPUT rows owned by `rank`: [rank*n/world, (rank+1)*n/world)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def global_id_map(local_ids, offset):
    """local_ids[i] = PointId of the shard's i-th row  ->  map[pid] = offset + i  (what idb_index_set_id_map takes)."""
    m = np.empty(len(local_ids), dtype=np.uint32)
    m[np.asarray(local_ids, dtype=np.int64)] = np.arange(offset, offset + len(local_ids), dtype=np.uint32)
    return m


def pack_keys(dist, gids, lens):
    """(distance bits << 32 | global id); slots beyond each query's result length are KEY_NONE."""
    d = np.ascontiguousarray(dist, dtype=np.float32).view(np.uint32).astype(np.uint64)
    keys = (d << np.uint64(32)) | np.asarray(gids, dtype=np.uint64)
    j = np.arange(keys.shape[1])[None, :]
    keys[j >= np.asarray(lens)[:, None]] = KEY_NONE
    return keys


def merge_keys(all_keys, k):
    """Host statement of the merge kernel: all_keys [world, nq, k] -> (ids [nq,k], dist [nq,k], lens [nq])."""
    w, nq, kk = all_keys.shape
    flat = np.sort(np.transpose(all_keys, (1, 0, 2)).reshape(nq, w * kk), axis=1)[:, :k]
    ids = (flat & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    dist = (flat >> np.uint64(32)).astype(np.uint32).view(np.float32).copy()
    none = flat == KEY_NONE
    ids[none] = 0xFFFFFFFF
    dist[none] = np.inf
    return ids, dist, (~none).sum(axis=1).astype(np.uint32)


class ShardedIndex:
    """One rank's shard + communicator.  `exchange_unique_id(bytes_or_None) -> bytes` broadcasts rank 0's NCCL id
    through whatever control plane the host application has (torch.distributed, MPI, a file ...)."""

    def __init__(self, rows_of_this_rank, offset, rank, world, device, exchange_unique_id, **build_kw):
        self.rank, self.world = rank, world
        self.index, local_ids = _abi.Index.build(rows_of_this_rank, device=device, **build_kw)
        self.index.set_id_map(global_id_map(local_ids, offset))
        uid = exchange_unique_id(_abi.comm_unique_id() if rank == 0 else None)
        self.comm = _abi.Comm(uid, rank, world, device)

    def search(self, queries, ef_search=0, k=10):
        return self.index.sharded_search(self.comm, queries, ef_search, k)

    def close(self):
        self.comm.close()
        self.index.close()
