"""Batched independent worlds (RL-style) on one GPU: thin mirror of the b2gpu_batch_* C ABI."""
import ctypes as C

import numpy as np

from . import abi
from .lib import check, load


class Context:
    """One per device (b2gpu_init).  `stream` may be a raw cudaStream_t (e.g. torch's)."""

    def __init__(self, device=0, stream=None, lib_path=None):
        self.L = load(lib_path)
        self.h = C.c_void_p()
        check(self.L, self.L.b2gpu_init(device, C.c_void_p(stream) if stream else None, C.byref(self.h)))

    def sync(self):
        check(self.L, self.L.b2gpu_sync(self.h))

    def launch_count(self):
        return int(self.L.b2gpu_launch_count(self.h))

    def stream(self):
        return self.L.b2gpu_stream(self.h)

    def set_profiling(self, on):
        check(self.L, self.L.b2gpu_set_profiling(self.h, int(on)))

    def stage_times(self):
        """{stage name: (milliseconds, launches)} accumulated since set_profiling(True)."""
        n = self.L.b2gpu_stage_count()
        ms = np.zeros(n, np.float64)
        cnt = np.zeros(n, np.int64)
        check(self.L, self.L.b2gpu_get_stage_times(self.h, ms.ctypes.data, cnt.ctypes.data, n))
        return {self.L.b2gpu_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def close(self):
        if self.h:
            self.L.b2gpu_shutdown(self.h)
            self.h = C.c_void_p()


class Batch:
    def __init__(self, ctx, proto, n_worlds, max_contacts=0, max_pairs=0, lane_block=0, generic_solver=False, solver=None):
        """proto: abi.Snapshot of the prototype world (from B2world.snapshot())."""
        self.ctx, self.L = ctx, ctx.L
        caps = abi.Caps()
        caps.max_contacts, caps.max_pairs = max_contacts, max_pairs
        caps.reserved[0] = lane_block
        # solver (diagnostic switch, b2gpu_caps.reserved[1]): None = default kernels (straight-line velocity and position,
        # 8 stream groups, CUDA graphs); 'generic' = global-memory stages; 'one_stream' = no stream groups; 'no_graph' = no
        # CUDA graphs; 'large' = large-world mode (exactly one world: data-parallel broadphase / islands, b2g_large.h);
        # 'large_exact' = large-world mode that keeps the replica tree (reference contact order, bit-identical free-running).
        # (Round 1's rejected kernel variants — pipelined, TMA ring, producer warp, level-scheduled — were removed.)
        codes = {None: 0, 'generic': 1, 'one_stream': 5, 'no_graph': 6, 'large': 11, 'large_exact': 12}
        caps.reserved[1] = 1 if generic_solver else codes[solver]
        self.h = C.c_void_p()
        c = proto.as_c()
        self._keep = proto
        check(self.L, self.L.b2gpu_batch_create(ctx.h, C.byref(c), n_worlds, C.byref(caps), C.byref(self.h)))
        self.n_worlds = n_worlds
        self.body_count = proto.n.body_count

    def close(self):
        if self.h:
            self.L.b2gpu_batch_destroy(self.h)
            self.h = C.c_void_p()

    def step(self, dt, velocity_iterations, position_iterations, steps=1):
        check(self.L, self.L.b2gpu_batch_step(self.h, dt, velocity_iterations, position_iterations, steps))

    def reset(self, snap):
        """Every world back to the state of `snap` (b2gpu_batch_reset: an RL-style reset of all environments)."""
        c = snap.as_c()
        check(self.L, self.L.b2gpu_batch_reset(self.h, C.byref(c)))

    def check_status(self):
        """Synchronises and raises B2gpuError if a world failed on the device (capacity overflow, unsupported shape
        pair).  step() is asynchronous and cannot report it; step_host / body_state / download_world do."""
        check(self.L, self.L.b2gpu_batch_status(self.h))

    def set_level_threshold(self, contacts):
        """Large-world batches: islands with at least `contacts` contacts (no joints) are swept in dependency-level order by a
        CTA (b2gpu_batch_set_level_threshold); 0 = library default (1024), negative = never."""
        check(self.L, self.L.b2gpu_batch_set_level_threshold(self.h, int(contacts)))

    def upload_world(self, world, snap):
        c = snap.as_c()
        check(self.L, self.L.b2gpu_batch_upload_world(self.h, world, C.byref(c)))

    def download_world(self, world):
        n = abi.SnapshotSizes()
        check(self.L, self.L.b2gpu_batch_snapshot_sizes(self.h, world, C.byref(n)))
        snap = abi.Snapshot(n)
        c = snap.as_c()
        check(self.L, self.L.b2gpu_batch_download_world(self.h, world, C.byref(c)))
        return snap.finish(c)

    def step_with_events(self, dt, velocity_iterations, position_iterations, worlds):
        """One step of the whole batch plus, for the listed worlds, the begin_contact / end_contact events of that step
        in the reference's firing order (b2gpu_contact_events): {world: abi.CONTACT_EVENT_DTYPE array}.  Costs two
        snapshot downloads per listed world — meant for the few worlds somebody is watching, not for all of them."""
        from .world import contact_events
        before = {w: self.download_world(w) for w in worlds}
        self.step(dt, velocity_iterations, position_iterations)
        return {w: contact_events(self.L, before[w], self.download_world(w), int(self.stats(w, 1)[0]["destroyed"])) for w in worlds}

    def post_solve_events(self, world):
        """post_solve reports of the last step of one world of the batch (b2gpu_batch_post_solve_events)."""
        n = check(self.L, self.L.b2gpu_batch_post_solve_events(self.h, world, None, 0))
        out = np.zeros(max(n, 1), abi.POST_SOLVE_DTYPE)
        check(self.L, self.L.b2gpu_batch_post_solve_events(self.h, world, out.ctypes.data, n))
        return out[:n]

    def save_checkpoint(self, world, path):
        """One world of the batch to a snapshot file (b2gpu_snapshot_save)."""
        from . import checkpoint
        checkpoint.save(self.download_world(world), path, self.L)

    def load_checkpoint(self, world, path):
        from . import checkpoint
        self.upload_world(world, checkpoint.load(path, self.L))

    def stats(self, first=0, count=None):
        count = self.n_worlds - first if count is None else count
        out = np.zeros(count, abi.STATS_DTYPE)
        check(self.L, self.L.b2gpu_batch_get_stats(self.h, first, count, out.ctypes.data))
        return out

    def body_state(self, first=0, count=None):
        count = self.n_worlds - first if count is None else count
        out = np.zeros((count, self.body_count, 8), np.float32)
        check(self.L, self.L.b2gpu_batch_get_body_state(self.h, out.ctypes.data, first, count))
        return out

    def step_host(self, forces, state_out, dt, velocity_iterations, position_iterations, steps=1):
        """End-to-end step through HOST buffers: H2D forces, step(s), D2H body state (synchronous)."""
        fp = forces.ctypes.data if forces is not None else None
        check(self.L, self.L.b2gpu_batch_step_host(self.h, fp, state_out.ctypes.data, dt, velocity_iterations,
                                                   position_iterations, steps))

    def device_buffers(self):
        """(forces pointer, forces bytes, state pointer, state bytes): raw device addresses of the staging buffers for zero-copy
        consumers (e.g. torch tensors); see apply_device_forces / refresh_device_state."""
        nf, ns = C.c_int64(), C.c_int64()
        pf = self.L.b2gpu_batch_forces_device(self.h, C.byref(nf))
        ps = self.L.b2gpu_batch_body_state_device(self.h, C.byref(ns))
        return pf, nf.value, ps, ns.value

    def apply_device_forces(self):
        check(self.L, self.L.b2gpu_batch_apply_device_forces(self.h))

    def refresh_device_state(self):
        check(self.L, self.L.b2gpu_batch_refresh_device_state(self.h))

    def dynamic_bodies(self):
        """Body indices of the prototype's dynamic bodies (the rows of the compact I/O arrays)."""
        n = check(self.L, self.L.b2gpu_batch_dynamic_bodies(self.h, None, 0))
        out = np.zeros(max(n, 1), np.int32)
        check(self.L, self.L.b2gpu_batch_dynamic_bodies(self.h, out.ctypes.data, n))
        return out[:n]

    def step_host_dynamic(self, forces, state_out, dt, velocity_iterations, position_iterations, steps=1):
        """step_host with compact I/O: forces [n_worlds][nd][3] (or None), state_out [n_worlds][nd][6] = c.x c.y a v.x v.y w
        of the dynamic bodies only."""
        fp = forces.ctypes.data if forces is not None else None
        check(self.L, self.L.b2gpu_batch_step_host_dynamic(self.h, fp, state_out.ctypes.data, dt, velocity_iterations,
                                                           position_iterations, steps))

    def ray_cast_closest(self, p1p2):
        """Closest-hit ray casts in every world: p1p2 [n_worlds][rays][4] -> abi.RAY_HIT_DTYPE [n_worlds][rays]."""
        rays = np.ascontiguousarray(p1p2, np.float32)
        assert rays.ndim == 3 and rays.shape[0] == self.n_worlds and rays.shape[2] == 4
        out = np.zeros(rays.shape[:2], abi.RAY_HIT_DTYPE)
        check(self.L, self.L.b2gpu_batch_ray_cast_closest(self.h, rays.ctypes.data, rays.shape[1], out.ctypes.data))
        return out

    def query_aabb(self, aabbs, max_hits=64):
        """B2world::query_aabb in every world: aabbs [n_worlds][boxes][4] -> (hits [n_worlds][boxes][max_hits][2] of
        (fixture, child) in report order, counts [n_worlds][boxes]; counts may exceed max_hits: hits are truncated, and
        entries past min(count, max_hits) of a box are unspecified)."""
        boxes = np.ascontiguousarray(aabbs, np.float32)
        assert boxes.ndim == 3 and boxes.shape[0] == self.n_worlds and boxes.shape[2] == 4
        counts = np.zeros(boxes.shape[:2], np.int32)
        hits = np.full(boxes.shape[:2] + (max(max_hits, 1), 2), -1, np.int32)
        check(self.L, self.L.b2gpu_batch_query_aabb(self.h, boxes.ctypes.data, boxes.shape[1], max_hits, counts.ctypes.data,
                                                     hits.ctypes.data))
        return hits, counts

    def algorithmic_bytes(self):
        return int(self.L.b2gpu_batch_algorithmic_bytes(self.h))

    def set_forces(self, forces, first=0):
        f = np.ascontiguousarray(forces, np.float32)
        assert f.shape[1:] == (self.body_count, 3)
        check(self.L, self.L.b2gpu_batch_set_forces(self.h, f.ctypes.data, first, f.shape[0]))

    def set_gravity(self, gxgy, first=0):
        """B2world::set_gravity per world: gxgy[n][2] (domain randomisation)."""
        v = np.ascontiguousarray(gxgy, np.float32).reshape(-1, 2)
        check(self.L, self.L.b2gpu_batch_set_gravity(self.h, v.ctypes.data, first, v.shape[0]))

    def _joint_control(self, joint, control, values, per, first):
        v = np.ascontiguousarray(values, np.float32).reshape(-1, per)
        check(self.L, self.L.b2gpu_batch_set_joint_control(self.h, getattr(joint, "index", joint), control, v.ctypes.data, first, v.shape[0]))

    def set_motor_speeds(self, joint, speeds, first=0):
        """B2revoluteJoint::set_motor_speed in worlds first .. first + len(speeds): one value per world (an RL action)."""
        self._joint_control(joint, abi.JOINT_CONTROL_MOTOR_SPEED, speeds, 1, first)

    def set_max_motor_torques(self, joint, torques, first=0):
        """B2revoluteJoint::set_max_motor_torque per world (prismatic joints: the maximum motor force)."""
        self._joint_control(joint, abi.JOINT_CONTROL_MAX_MOTOR_TORQUE, torques, 1, first)

    def set_targets(self, joint, targets, first=0):
        """B2mouseJoint::set_target per world: targets[n][2]."""
        self._joint_control(joint, abi.JOINT_CONTROL_TARGET, targets, 2, first)

    def set_linear_velocity(self, body, vxvy, first=0):
        v = np.ascontiguousarray(vxvy, np.float32)
        check(self.L, self.L.b2gpu_batch_set_linear_velocity(self.h, body, v.ctypes.data, first, v.shape[0]))
