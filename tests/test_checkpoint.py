"""Checkpoint / resume (second half): the snapshot file of b2gpu_snapshot_save / _load.

CPU tests drive the file calls of the product library (host-only code, no device needed) with oracle snapshots of
every scene family, resume in the test-only host simulator, and check the rejection paths (corrupt, truncated,
foreign, too-small arrays, out-of-range indices).  The GPU test saves from the device mid-run, resumes in a fresh
world and in a batch, and compares with the uninterrupted oracle run bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
from conftest import HOSTSIM_SO, SCENES


def _oracle(name, steps):
    from box2d_rs_b200 import scenes
    from oracle import b2o
    recipe, gravity, _ = SCENES[name]
    wo = b2o.B2world(gravity)
    recipe(scenes, wo)
    for _ in range(steps):
        wo.step(scenes.DT, 8, 3)
    return wo


def _same_bytes(a, b):
    assert bytes(a.world) == bytes(b.world) and bytes(a.n) == bytes(b.n)
    for f in ("bodies", "fixtures", "shapes", "proxies", "nodes", "contacts", "move_buffer"):
        assert getattr(a, f).tobytes() == getattr(b, f).tobytes(), f


@pytest.mark.parametrize("name", list(SCENES))
def test_file_round_trip_every_scene(name, built, tmp_path):
    """save -> load returns every byte of every table, at t = 0 (move buffer full, no contacts) and mid-run."""
    from box2d_rs_b200 import checkpoint
    for steps in (0, 25):
        snap = _oracle(name, steps).snapshot()
        checkpoint.validate(snap)
        path = tmp_path / ("%s_%d.b2snap" % (name, steps))
        checkpoint.save(snap, path)
        n = checkpoint.file_sizes(path)
        assert bytes(n) == bytes(snap.n)
        back = checkpoint.load(path)
        _same_bytes(snap, back)
        assert parity.compare_snapshots(snap, back) == []
        assert not os.path.exists(str(path) + ".tmp")


@pytest.mark.parametrize("name", ["pyramid", "mixed300", "sensors", "terrain"])
def test_resume_in_host_simulator(name, built, tmp_path):
    """A world restored from a file continues the saved run: oracle 40 steps -> file -> simulator +30 steps ==
    oracle 70 steps, every field; and a file written by the simulator restores the oracle's state."""
    from box2d_rs_b200 import batch, checkpoint, scenes, world
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    wo = _oracle(name, 40)
    path = tmp_path / "mid.b2snap"
    checkpoint.save(wo.snapshot(), path)
    wg = world.B2world(SCENES[name][1], ctx=ctx)
    wg.load_checkpoint(path)
    for _ in range(30):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
    assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
    path2 = tmp_path / "late.b2snap"
    wg.save_checkpoint(path2)
    assert parity.compare_snapshots(wo.snapshot(), checkpoint.load(path2)) == []
    # one world of a batch, too
    b = wg.batch(3, lane_block=1)
    b.load_checkpoint(1, path)
    for _ in range(30):
        b.step(scenes.DT, 8, 3)
    assert parity.compare_snapshots(wo.snapshot(), b.download_world(1)) == []
    b.close()
    wg.close()
    ctx.close()


def test_rejects_bad_files(built, tmp_path):
    from box2d_rs_b200 import abi, checkpoint, lib
    snap = _oracle("hello_world", 10).snapshot()
    good = tmp_path / "good.b2snap"
    checkpoint.save(snap, good)
    raw = good.read_bytes()

    def expect(data, code, name):
        p = tmp_path / name
        p.write_bytes(data)
        with pytest.raises(lib.B2gpuError) as e:
            checkpoint.load(p)
        assert e.value.code == code, (name, str(e.value))

    expect(b"", abi.E_INVALID, "empty")
    expect(b"not a snapshot" * 20, abi.E_INVALID, "foreign")
    expect(raw[:-5], abi.E_INVALID, "truncated")
    expect(raw + b"\0", abi.E_INVALID, "trailing")
    flipped = bytearray(raw)
    flipped[len(raw) // 2] ^= 0x10
    expect(bytes(flipped), abi.E_INVALID, "payload_bit_flip")
    flipped = bytearray(raw)
    flipped[40] ^= 0x01  # a record size in the header
    expect(bytes(flipped), abi.E_INVALID, "header_bit_flip")
    with pytest.raises(lib.B2gpuError) as e:
        checkpoint.load(tmp_path / "does_not_exist")
    assert e.value.code == abi.E_IO
    with pytest.raises(lib.B2gpuError) as e:
        checkpoint.save(snap, tmp_path / "no_such_dir" / "x.b2snap")
    assert e.value.code == abi.E_IO
    # caller arrays smaller than the file's tables
    L = lib.load()
    small = abi.Snapshot(abi.SnapshotSizes(1, 1, 1, 1, 1, 1, 1, 0))
    c = small.as_c()
    assert L.b2gpu_snapshot_load(os.fsencode(good), C.byref(c)) == abi.E_CAPACITY
    assert small.bodies[0]["type"] == 0 and c.n.body_count == 1  # untouched
    # big enough capacities but a NULL array: an error code, not a crash
    roomy = abi.Snapshot(checkpoint.file_sizes(good))
    c = roomy.as_c()
    c.bodies = None
    assert L.b2gpu_snapshot_load(os.fsencode(good), C.byref(c)) == abi.E_INVALID


def test_validate_catches_out_of_range_indices(built, tmp_path):
    from box2d_rs_b200 import abi, checkpoint, lib
    base = _oracle("pyramid", 30).snapshot()
    assert len(base.contacts) > 100

    def broken(mutate):
        s = _oracle("pyramid", 30).snapshot()
        mutate(s)
        with pytest.raises(lib.B2gpuError) as e:
            checkpoint.validate(s)
        assert e.value.code == abi.E_INVALID
        with pytest.raises(lib.B2gpuError):
            checkpoint.save(s, tmp_path / "bad.b2snap")
        assert not os.path.exists(tmp_path / "bad.b2snap")

    def set_field(table, i, field, value):
        def m(s):
            getattr(s, table)[i][field] = value
        return m

    broken(set_field("contacts", 7, "fixture_b", len(base.fixtures)))
    broken(set_field("contacts", 3, "index_a", 5))
    broken(set_field("fixtures", 2, "body", -1))
    broken(set_field("fixtures", 2, "shape_first", len(base.shapes)))
    broken(set_field("proxies", 4, "proxy_id", len(base.nodes)))
    broken(set_field("nodes", 9, "child1", len(base.nodes) + 3))
    broken(set_field("bodies", 1, "fixture_head", 10 ** 6))

    def bad_root(s):
        s.world.tree_root = len(s.nodes)
    broken(bad_root)

    def bad_manifold(s):
        s.contacts[0]["manifold"]["point_count"] = 3
    broken(bad_manifold)

    # structural damage that keeps every index in range: the step would loop forever or read link[-1]
    broken(set_field("fixtures", 0, "next", 0))                       # fixture list cycle
    broken(set_field("fixtures", 5, "body", 3))                       # fixture on the list of another body
    broken(set_field("bodies", 4, "fixture_count", 2))                # count does not match the list

    def internal(s):
        return int(np.flatnonzero(s.nodes["height"] > 0)[3])

    def leaf(s):
        return int(np.flatnonzero((s.nodes["height"] == 0) & (s.nodes["proxy"] >= 0))[3])

    def half_linked(s):
        s.nodes[internal(s)]["child2"] = -1
    broken(half_linked)

    def self_cycle(s):
        i = internal(s)
        s.nodes[i]["child1"] = i
    broken(self_cycle)

    def leaf_without_proxy(s):
        s.nodes[leaf(s)]["proxy"] = -1
    broken(leaf_without_proxy)

    def wrong_back_link(s):
        i = internal(s)
        s.nodes[int(s.nodes[i]["child1"])]["parent"] = -1
    broken(wrong_back_link)

    def move_buffer_names_internal_node(s):
        assert len(s.move_buffer) > 0
        s.move_buffer[0] = internal(s)

    def with_moves(mutate):
        o = _oracle("pyramid", 30)
        o.body(211).set_transform((0.0, 30.0), 0.1)  # buffers a move
        s = o.snapshot()
        mutate(s)
        with pytest.raises(lib.B2gpuError) as e:
            checkpoint.validate(s)
        assert e.value.code == abi.E_INVALID
    with_moves(move_buffer_names_internal_node)

    def bad_node_count(s):
        s.world.tree_node_count += 1
    broken(bad_node_count)

    def proxy_of_other_fixture(s):
        s.proxies[6]["fixture"] = 3
    broken(proxy_of_other_fixture)


@pytest.mark.gpu
def test_gpu_resume_is_bit_identical(built, tmp_path):
    """Device run saved mid-flight, resumed in a fresh world and in one world of a batch == uninterrupted oracle."""
    from box2d_rs_b200 import checkpoint, scenes, world
    for name in ("pyramid", "mixed300"):
        recipe, gravity, _ = SCENES[name]
        wo = _oracle(name, 0)
        wg = world.B2world(gravity)
        recipe(scenes, wg)
        for _ in range(50):
            wo.step(scenes.DT, 8, 3)
            wg.step(scenes.DT, 8, 3)
        path = tmp_path / (name + ".b2snap")
        wg.save_checkpoint(path)
        assert parity.compare_snapshots(wo.snapshot(), checkpoint.load(path)) == []
        w2 = world.B2world(gravity, ctx=wg.ctx)
        w2.load_checkpoint(path)
        b = w2.batch(33)
        for _ in range(40):
            wo.step(scenes.DT, 8, 3)
            w2.step(scenes.DT, 8, 3)
            b.step(scenes.DT, 8, 3)
        assert parity.compare_snapshots(wo.snapshot(), w2.snapshot()) == []
        assert parity.compare_snapshots(wo.snapshot(), b.download_world(32)) == []
        b.close()
        w2.close()
        wg.close()


def test_committed_file_still_loads(built):
    """Format stability: the version-1 file committed under tests/golden/ (make_checkpoint_golden.py) loads, agrees
    with the golden fixture of the same scene at step 30, and resumed for 30 steps reaches its step-60 record."""
    import sys
    from conftest import ROOT
    from box2d_rs_b200 import batch, checkpoint, scenes, world
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    g = np.load(os.path.join(ROOT, "tests", "golden", "hello_world.npz"))
    snap = checkpoint.load(os.path.join(ROOT, "tests", "golden", "hello_world_step30.b2snap"))
    assert np.array_equal(make_golden.contact_table(snap), g["contacts_30"])
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    wg.upload(snap)
    b = wg.batch(1, lane_block=1)
    assert np.array_equal(b.body_state()[0].view(np.uint32), g["state_30"].view(np.uint32))
    b.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, steps=30)
    assert np.array_equal(b.body_state()[0].view(np.uint32), g["state_60"].view(np.uint32))
    assert np.array_equal(make_golden.contact_table(b.download_world(0)), g["contacts_60"])
    b.close()
    wg.close()
    ctx.close()


def test_sharded_batch_resumes_under_another_world_size(built, tmp_path):
    """7 perturbed Pyramid worlds stepped as 2 shards, saved per global world index, resumed as 3 shards: every
    world ends where the same world of one unsharded batch ends (worlds are the only unit that shards)."""
    from box2d_rs_b200 import batch, checkpoint, scenes, sharding, world
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    scenes.pyramid(wg)
    total, seed = 7, 2864

    def make(first, end):
        b = wg.batch(end - first, lane_block=2, max_contacts=800)
        v = sharding.perturbation(first, end - first, seed)
        b.set_linear_velocity(211, v)
        return b

    whole = make(0, total)
    whole.step(scenes.DT, 8, 3, steps=50)
    for rank in range(2):
        first, end = sharding.world_range(total, rank, 2)
        b = make(first, end)
        b.step(scenes.DT, 8, 3, steps=20)
        assert checkpoint.save_batch(b, tmp_path / "ckpt", first) == end - first
        b.close()
    assert len(os.listdir(tmp_path / "ckpt")) == total
    for rank in range(3):
        first, end = sharding.world_range(total, rank, 3)
        b = wg.batch(end - first, lane_block=1, max_contacts=800)
        checkpoint.load_batch(b, tmp_path / "ckpt", first)
        b.step(scenes.DT, 8, 3, steps=30)
        for w in range(end - first):
            assert parity.compare_snapshots(whole.download_world(first + w), b.download_world(w)) == [], (rank, w)
        b.close()
    whole.close()
    wg.close()
    ctx.close()
