"""Checkpoint / resume: abi.Snapshot <-> the snapshot file of b2gpu_snapshot_save / b2gpu_snapshot_load
(box2d_rs_b200/csrc/b2g_checkpoint.cu).  The reference serialises world *definitions* with serde
(src/serialize/serialize_b2_world.rs:133-178); a snapshot file carries the whole step state (contacts, impulses,
tree, move buffer), so `load` + `upload` + `step` continues a run bit for bit.  Host-only: no device needed."""
import ctypes as C
import os

from . import abi
from .lib import check, load as _load_lib


def validate(snap, L=None):
    """Raise B2gpuError(E_INVALID) unless every index of the snapshot stays inside its table."""
    L = L or _load_lib()
    c = snap.as_c()
    check(L, L.b2gpu_snapshot_validate(C.byref(c)))


def save(snap, path, L=None):
    """Write an abi.Snapshot to `path` (atomically: `path`.tmp, then rename)."""
    L = L or _load_lib()
    c = snap.as_c()
    check(L, L.b2gpu_snapshot_save(C.byref(c), os.fsencode(path)))


def file_sizes(path, L=None):
    L = L or _load_lib()
    n = abi.SnapshotSizes()
    check(L, L.b2gpu_snapshot_file_sizes(os.fsencode(path), C.byref(n)))
    return n


def load(path, L=None):
    """Read a snapshot file into a new abi.Snapshot (ready for B2world.upload / Batch.upload_world)."""
    L = L or _load_lib()
    snap = abi.Snapshot(file_sizes(path, L))
    c = snap.as_c()
    check(L, L.b2gpu_snapshot_load(os.fsencode(path), C.byref(c)))
    return snap.finish(c)


# ---- batches: one snapshot file per world, named by GLOBAL world index, so a run sharded over N ranks
# (sharding.world_range) can be resumed over M ranks: each rank reads the files of the worlds it owns now.
def world_path(directory, global_world):
    return os.path.join(os.fspath(directory), "world_%08d.b2snap" % global_world)


def save_batch(batch, directory, first_global_world=0):
    """Every world of this rank's batch to `directory`; `first_global_world` = sharding.world_range(...)[0].
    Ranks write disjoint files, so no coordination is needed beyond a barrier before anybody reads."""
    os.makedirs(directory, exist_ok=True)
    for w in range(batch.n_worlds):
        save(batch.download_world(w), world_path(directory, first_global_world + w), batch.L)
    return batch.n_worlds


def load_batch(batch, directory, first_global_world=0):
    """Restore every world of this rank's batch from the files of its global world indices."""
    for w in range(batch.n_worlds):
        batch.upload_world(w, load(world_path(directory, first_global_world + w), batch.L))
    return batch.n_worlds
