"""Snapshot comparison helpers shared by the CPU (host simulator) and GPU parity tests."""
import numpy as np

INT_BODY = ("type", "flags", "fixture_head", "fixture_count")
F_BODY = ("xf", "lc", "c0", "c", "a0", "a", "v", "w", "force", "torque", "mass", "inv_mass", "inertia", "inv_inertia",
          "linear_damping", "angular_damping", "gravity_scale", "sleep_time")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def fdiff(a, b, rtol, atol):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b))


def compare_snapshots(ref, got, rtol=0.0, atol=0.0, what="", check_tree=True, body_flag_mask=0xFFFF):
    """ref = oracle, got = engine.  Integer fields must be equal; floats bit-equal when rtol == atol == 0,
    else within rtol relative / atol absolute.  Returns a list of mismatch strings (empty = parity)."""
    bad = []

    def chk_int(name, a, b):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            bad.append("%s%s: shape %s vs %s" % (what, name, a.shape, b.shape))
        elif not np.array_equal(a, b):
            idx = np.argwhere(a != b)[:3].tolist()
            bad.append("%s%s: %d integer mismatches, first at %s" % (what, name, int((a != b).sum()), idx))

    def chk_f(name, a, b):
        a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
        if a.shape != b.shape:
            bad.append("%s%s: shape %s vs %s" % (what, name, a.shape, b.shape))
            return
        if rtol == 0.0 and atol == 0.0:
            ok = _bits(a) == _bits(b)
        else:
            ok = fdiff(a, b, rtol, atol)
        if not ok.all():
            idx = np.argwhere(~ok)[:3].tolist()
            i0 = tuple(idx[0])
            bad.append("%s%s: %d float mismatches, first at %s: ref %r got %r" %
                       (what, name, int((~ok).sum()), idx, float(a[i0]), float(b[i0])))

    for f in ("body_count", "fixture_count", "shape_count", "proxy_count", "node_count", "contact_count", "move_count", "joint_count"):
        if getattr(ref.n, f) != getattr(got.n, f):
            bad.append("%ssizes.%s: %d vs %d" % (what, f, getattr(ref.n, f), getattr(got.n, f)))
    if bad:
        return bad
    for f in ("flags", "tree_root", "tree_free_list", "tree_node_count", "tree_node_capacity", "tree_insertion_count",
              "proxy_count"):
        if not check_tree and f.startswith("tree_"):
            continue
        if getattr(ref.world, f) != getattr(got.world, f):
            bad.append("%sworld.%s: %d vs %d" % (what, f, getattr(ref.world, f), getattr(got.world, f)))
    chk_f("world.inv_dt0", [ref.world.inv_dt0], [got.world.inv_dt0])
    chk_int("bodies.type", ref.bodies["type"], got.bodies["type"])
    chk_int("bodies.flags", ref.bodies["flags"] & body_flag_mask, got.bodies["flags"] & body_flag_mask)
    for f in F_BODY:
        chk_f("bodies." + f, ref.bodies[f], got.bodies[f])
    chk_f("proxies.aabb", ref.proxies["aabb"], got.proxies["aabb"])
    chk_int("proxies.ids", ref.proxies[["fixture", "child_index", "proxy_id"]].tolist(),
            got.proxies[["fixture", "child_index", "proxy_id"]].tolist())
    if check_tree:
        live = ref.nodes["height"] >= 0
        chk_int("nodes.height", ref.nodes["height"], got.nodes["height"])
        chk_int("nodes.parent", ref.nodes["parent"], got.nodes["parent"])
        for f in ("child1", "child2", "moved", "proxy"):
            chk_int("nodes." + f, ref.nodes[f][live], got.nodes[f][live])
        chk_f("nodes.aabb", ref.nodes["aabb"][live], got.nodes["aabb"][live])
    chk_int("move_buffer", ref.move_buffer, got.move_buffer)
    for f in ("type", "body_a", "body_b", "flags"):
        chk_int("joints." + f, ref.joints[f], got.joints[f])
    for f in ("local_anchor_a", "local_anchor_b", "param", "impulse"):
        chk_f("joints." + f, ref.joints[f], got.joints[f])
    for f in ("fixture_a", "fixture_b", "index_a", "index_b", "flags"):
        chk_int("contacts." + f, ref.contacts[f], got.contacts[f])
    for f in ("friction", "restitution", "restitution_threshold", "tangent_speed"):
        chk_f("contacts." + f, ref.contacts[f], got.contacts[f])
    rm, gm = ref.contacts["manifold"], got.contacts["manifold"]
    chk_int("manifold.type", rm["type"], gm["type"])
    chk_int("manifold.point_count", rm["point_count"], gm["point_count"])
    chk_int("manifold.ids", rm["points"]["id"], gm["points"]["id"])
    for f in ("ln", "lp"):
        chk_f("manifold." + f, rm[f], gm[f])
    for f in ("lp", "normal_impulse", "tangent_impulse"):
        chk_f("manifold.points." + f, rm["points"][f], gm["points"][f])
    return bad


STAT_FIELDS = ("status", "contacts", "touching", "destroyed", "islands", "island_bodies", "island_contacts", "moved", "pairs",
               "created", "awake_bodies")


def compare_stats(ref, got, what=""):
    return ["%sstats.%s: %d vs %d" % (what, f, int(ref[f]), int(got[f])) for f in STAT_FIELDS if int(ref[f]) != int(got[f])]


def reorder_created(ref, got, created):
    """Large-world mode appends the contacts created by one update_pairs call in LBVH order instead of
    reference-tree order: permute the last `created` contacts of `got` into `ref`'s order, matching them by
    (fixture_a, fixture_b, index_a, index_b).  Returns a mismatch string when the created SETS differ."""
    n = ref.n.contact_count
    if created == 0 or n != got.n.contact_count:
        return None
    t0 = n - created

    def key(c):
        return int(c["fixture_a"]), int(c["fixture_b"]), int(c["index_a"]), int(c["index_b"])
    pos = {key(got.contacts[i]): i for i in range(t0, n)}
    if len(pos) != created:
        return "created contacts: duplicate fixture pair"
    try:
        perm = [pos[key(ref.contacts[i])] for i in range(t0, n)]
    except KeyError as e:
        return "created contact set differs: %s missing" % (e,)
    got.contacts[t0:n] = got.contacts[perm]
    return None


def compare_large_step(ref, got, ref_stats, got_stats):
    """One teacher-forced step of the large-world mode against the oracle: everything bit-equal except the
    order of the contacts created in this step (compared as a set), the replica tree (not maintained) and the
    island_bodies counter (static bodies are not listed in island arrays)."""
    msg = reorder_created(ref, got, int(ref_stats["created"]))
    bad = ([msg] if msg else []) + compare_snapshots(ref, got, check_tree=False)
    return bad + [b for b in compare_stats(ref_stats, got_stats) if "island_bodies" not in b]
