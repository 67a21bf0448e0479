//! examples/dump_state.rs — the oracle pinning kit.
//!
//! Runs the scene recipes of box2d_rs_b200/scenes.py on the REAL box2d-rs crate and writes the full step state after
//! chosen steps as b2gpu snapshot files (`Snapshot::save` of src/private/gpu/mirror.rs — pure Rust, no GPU, no nvcc):
//!
//!     cargo run --release --example dump_state -- /path/to/b200-repo/tests/reference_dump
//!     python -m pytest tests/test_reference_dump.py            # compares every file with the C++ oracle, bit for bit
//!
//! NOT BUILT IN THIS REPOSITORY (no Rust toolchain in the image).  Needs `B2world::gpu_snapshot`.
//! The recipes mirror scenes.py line for line: positions are computed in f64 and rounded once to f32 (`as f32`), the
//! random streams are SplitMix64 with the same seeds, bodies get their fixtures before the next body is created.
use std::cell::RefCell;
use std::path::PathBuf;
use std::rc::Rc;

use box2d_rs::b2_body::*;
use box2d_rs::b2_fixture::*;
use box2d_rs::b2_joint::*;
use box2d_rs::b2_math::*;
use box2d_rs::b2_world::*;
use box2d_rs::b2rs_common::UserDataType;
use box2d_rs::joints::b2_distance_joint::*;
use box2d_rs::joints::b2_gear_joint::*;
use box2d_rs::joints::b2_mouse_joint::*;
use box2d_rs::joints::b2_prismatic_joint::*;
use box2d_rs::joints::b2_pulley_joint::*;
use box2d_rs::joints::b2_revolute_joint::*;
use box2d_rs::shapes::b2_circle_shape::*;
use box2d_rs::shapes::b2_edge_shape::*;
use box2d_rs::shapes::b2_polygon_shape::*;

#[derive(Default, Copy, Clone, Debug, PartialEq)]
struct Ud;
impl UserDataType for Ud {
    type Fixture = i32;
    type Body = i32;
    type Joint = i32;
}
type World = B2worldPtr<Ud>;
type Body = BodyPtr<Ud>;

struct SplitMix64(u64);
impl SplitMix64 {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
    fn uniform(&mut self, lo: f64, hi: f64) -> f64 { lo + (hi - lo) * ((self.next() >> 40) as f64 / (1u64 << 24) as f64) }
}

fn body(world: &World, dynamic: bool, x: f64, y: f64, angle: f64) -> Body {
    let mut bd = B2bodyDef::default();
    if dynamic { bd.body_type = B2bodyType::B2DynamicBody; }
    bd.position.set(x as f32, y as f32);
    bd.angle = angle as f32;
    B2world::create_body(world.clone(), &bd)
}
fn fixture(b: &Body, shape: Rc<RefCell<dyn box2d_rs::b2_shape::B2shapeDynTrait>>, density: f32, friction: f32) {
    let mut fd = B2fixtureDef::default();
    fd.shape = Some(shape);
    fd.density = density;
    fd.friction = friction;
    B2body::create_fixture(b.clone(), &fd);
}
fn boxed(hx: f32, hy: f32) -> Rc<RefCell<B2polygonShape>> {
    let mut s = B2polygonShape::default();
    s.set_as_box(hx, hy);
    Rc::new(RefCell::new(s))
}
fn circle(r: f32) -> Rc<RefCell<B2circleShape>> {
    let mut s = B2circleShape::default();
    s.base.m_radius = r;
    Rc::new(RefCell::new(s))
}
fn edge(world: &World, b: &Body, x1: f32, y1: f32, x2: f32, y2: f32) {
    let _ = world;
    let mut s = B2edgeShape::default();
    s.set_two_sided(B2vec2::new(x1, y1), B2vec2::new(x2, y2));
    B2body::create_fixture_by_shape(b.clone(), Rc::new(RefCell::new(s)), 0.0);
}
fn container(world: &World, hw: f32, height: f32) -> Body {
    let g = body(world, false, 0.0, 0.0, 0.0);
    edge(world, &g, -hw, 0.0, hw, 0.0);
    edge(world, &g, -hw, 0.0, -hw, height);
    edge(world, &g, hw, 0.0, hw, height);
    g
}

// ---- scenes.py: hello_world, pyramid, pile, add_pair, bridge, tumbler (gears and pulleys further down)
fn hello_world(world: &World) {
    let g = body(world, false, 0.0, -10.0, 0.0);
    B2body::create_fixture_by_shape(g, boxed(50.0, 10.0), 0.0);
    let b = body(world, true, 0.0, 4.0, 0.0);
    fixture(&b, boxed(1.0, 1.0), 1.0, 0.3);
}
fn pyramid(world: &World) {
    body(world, false, 0.0, 0.0, 0.0); // the testbed's empty ground body (examples/testbed/test.rs:181-182)
    let g = body(world, false, 0.0, 0.0, 0.0);
    edge(world, &g, -40.0, 0.0, 40.0, 0.0);
    let shape = boxed(0.5, 0.5);
    let mut x = B2vec2::new(-7.0, 0.75);
    let dx = B2vec2::new(0.5625, 1.25);
    let dy = B2vec2::new(1.125, 0.0);
    for i in 0..20 {
        let mut y = x;
        for _ in i..20 {
            let b = body(world, true, y.x as f64, y.y as f64, 0.0);
            B2body::create_fixture_by_shape(b, shape.clone(), 5.0);
            y += dy;
        }
        x += dx;
    }
}
fn pile(world: &World, n: usize, width: f64) {
    let mut rng = SplitMix64(0xB2D + 4);
    container(world, (width / 2.0) as f32, 100.0);
    let (c, b) = (circle(0.125), boxed(0.125, 0.125));
    let pitch = 0.3f64;
    let cols = std::cmp::max(1, ((width - 2.0) / pitch) as usize);
    for k in 0..n {
        let (col, row) = (k % cols, k / cols);
        let px = -width / 2.0 + 1.0 + pitch * col as f64 + rng.uniform(-0.02, 0.02);
        let py = 0.3 + pitch * row as f64 + rng.uniform(-0.02, 0.02);
        let bd = body(world, true, px, py, 0.0);
        if k % 2 == 0 { fixture(&bd, c.clone(), 1.0, 0.1); } else { fixture(&bd, b.clone(), 1.0, 0.3); }
    }
}
fn add_pair(world: &World, n: usize) {
    let mut rng = SplitMix64(0xB2D + 5);
    let c = circle(0.1);
    for _ in 0..n {
        let (px, py) = (rng.uniform(-60.0, 0.0), rng.uniform(-10.0, 20.0));
        let b = body(world, true, px, py, 0.0);
        B2body::create_fixture_by_shape(b, c.clone(), 0.01);
    }
    let mut bd = B2bodyDef::default();
    bd.body_type = B2bodyType::B2DynamicBody;
    bd.position.set(-100.0, 5.0);
    bd.bullet = true;
    let b = B2world::create_body(world.clone(), &bd);
    B2body::create_fixture_by_shape(b.clone(), boxed(1.5, 1.5), 1.0);
    b.borrow_mut().set_linear_velocity(B2vec2::new(100.0, 0.0));
}
fn bridge(world: &World) {
    body(world, false, 0.0, 0.0, 0.0);
    let g = body(world, false, 0.0, 0.0, 0.0);
    edge(world, &g, -40.0, 0.0, 40.0, 0.0);
    let plank = boxed(0.5, 0.125);
    let mut prev = g.clone();
    let count = 30;
    for i in 0..count {
        let b = body(world, true, -14.5 + 1.0 * i as f64, 5.0, 0.0);
        fixture(&b, plank.clone(), 20.0, 0.2);
        let mut jd = B2revoluteJointDef::default();
        jd.initialize(prev.clone(), b.clone(), B2vec2::new((-15.0 + 1.0 * i as f64) as f32, 5.0));
        world.borrow_mut().create_joint(&B2JointDefEnum::RevoluteJoint(jd));
        prev = b;
    }
    let mut jd = B2revoluteJointDef::default();
    jd.initialize(prev, g, B2vec2::new((-15.0 + 1.0 * count as f64) as f32, 5.0));
    world.borrow_mut().create_joint(&B2JointDefEnum::RevoluteJoint(jd));
    let mut tri = B2polygonShape::default();
    tri.set(&[B2vec2::new(-0.5, 0.0), B2vec2::new(0.5, 0.0), B2vec2::new(0.0, 1.5)]);
    let tri = Rc::new(RefCell::new(tri));
    for i in 0..2 {
        let b = body(world, true, -8.0 + 8.0 * i as f64, 12.0, 0.0);
        fixture(&b, tri.clone(), 1.0, 0.2);
    }
    let ball = circle(0.5);
    for i in 0..3 {
        let b = body(world, true, -6.0 + 6.0 * i as f64, 10.0, 0.0);
        fixture(&b, ball.clone(), 1.0, 0.2);
    }
}
fn tumbler(world: &World, n: usize) {
    let mut rng = SplitMix64(0xB2D + 21);
    let g = body(world, false, 0.0, 0.0, 0.0);
    let mut bd = B2bodyDef::default();
    bd.body_type = B2bodyType::B2DynamicBody;
    bd.allow_sleep = false;
    bd.position.set(0.0, 10.0);
    let drum = B2world::create_body(world.clone(), &bd);
    for (hx, hy, cx, cy) in [(0.5f32, 10.0f32, 10.0f32, 0.0f32), (0.5, 10.0, -10.0, 0.0), (10.0, 0.5, 0.0, 10.0), (10.0, 0.5, 0.0, -10.0)] {
        let mut s = B2polygonShape::default();
        s.set_as_box_angle(hx, hy, B2vec2::new(cx, cy), 0.0);
        B2body::create_fixture_by_shape(drum.clone(), Rc::new(RefCell::new(s)), 5.0);
    }
    let mut jd = B2revoluteJointDef::default();
    jd.base.body_a = Some(g);
    jd.base.body_b = Some(drum);
    jd.local_anchor_a.set(0.0, 10.0);
    jd.local_anchor_b.set(0.0, 0.0);
    jd.reference_angle = 0.0;
    jd.motor_speed = 0.05f32 * std::f32::consts::PI;
    jd.max_motor_torque = 1e8;
    jd.enable_motor = true;
    world.borrow_mut().create_joint(&B2JointDefEnum::RevoluteJoint(jd));
    let small = boxed(0.125, 0.125);
    for k in 0..n {
        let px = -4.0 + 0.4 * (k % 20) as f64 + rng.uniform(-0.05, 0.05);
        let py = 3.0 + 0.4 * (k / 20) as f64 + rng.uniform(-0.05, 0.05);
        let b = body(world, true, px, py, 0.0);
        B2body::create_fixture_by_shape(b, small.clone(), 1.0);
    }
}
fn pendulum(world: &World) {
    // a rigid distance joint (tests/test_joints.py::test_oracle_distance_joint_keeps_its_length)
    let g = body(world, false, 0.0, 0.0, 0.0);
    let bob = body(world, true, 3.0, 5.0, 0.0);
    B2body::create_fixture_by_shape(bob.clone(), circle(0.5), 1.0);
    let mut jd = B2distanceJointDef::default();
    jd.initialize(g, bob, B2vec2::new(0.0, 5.0), B2vec2::new(3.0, 5.0));
    world.borrow_mut().create_joint(&B2JointDefEnum::DistanceJoint(jd));
}

fn circle_at(r: f32, x: f32, y: f32) -> Rc<RefCell<B2circleShape>> {
    let mut s = B2circleShape::default();
    s.base.m_radius = r;
    s.m_p.set(x, y);
    Rc::new(RefCell::new(s))
}
fn revolute(world: &World, a: &Body, b: &Body, x: f32, y: f32) -> B2jointPtr<Ud> {
    let mut jd = B2revoluteJointDef::default();
    jd.initialize(a.clone(), b.clone(), B2vec2::new(x, y));
    world.borrow_mut().create_joint(&B2JointDefEnum::RevoluteJoint(jd))
}
fn gear(world: &World, a: &Body, b: &Body, j1: &B2jointPtr<Ud>, j2: &B2jointPtr<Ud>, ratio: f32) {
    let mut jd = B2gearJointDef::default();
    jd.base.body_a = Some(a.clone());
    jd.base.body_b = Some(b.clone());
    jd.joint1 = Some(j1.clone());
    jd.joint2 = Some(j2.clone());
    jd.ratio = ratio;
    world.borrow_mut().create_joint(&B2JointDefEnum::GearJoint(jd));
}
fn gears(world: &World) {
    // scenes.py::gears — the testbed's GearJoint scene plus two kicks
    let g = body(world, false, 0.0, 0.0, 0.0);
    edge(world, &g, 50.0, 0.0, -50.0, 0.0);
    let (circle1, circle2, bar) = (circle(1.0), circle(2.0), boxed(0.5, 5.0));
    let body1 = body(world, false, 10.0, 9.0, 0.0);
    B2body::create_fixture_by_shape(body1.clone(), circle1.clone(), 5.0);
    let body2 = body(world, true, 10.0, 8.0, 0.0);
    B2body::create_fixture_by_shape(body2.clone(), bar.clone(), 5.0);
    let body3 = body(world, true, 10.0, 6.0, 0.0);
    B2body::create_fixture_by_shape(body3.clone(), circle2.clone(), 5.0);
    let joint1 = revolute(world, &body1, &body2, 10.0, 9.0);
    let joint2 = revolute(world, &body2, &body3, 10.0, 6.0);
    gear(world, &body1, &body3, &joint1, &joint2, 2.0); // the STATIC disc as body A, as the testbed does
    body2.borrow_mut().set_angular_velocity(1.5);
    let b1 = body(world, true, -3.0, 12.0, 0.0);
    B2body::create_fixture_by_shape(b1.clone(), circle1, 5.0);
    let j1 = revolute(world, &g, &b1, -3.0, 12.0);
    let b2 = body(world, true, 0.0, 12.0, 0.0);
    B2body::create_fixture_by_shape(b2.clone(), circle2, 5.0);
    let j2 = revolute(world, &g, &b2, 0.0, 12.0);
    let b3 = body(world, true, 2.5, 12.0, 0.0);
    B2body::create_fixture_by_shape(b3.clone(), bar, 5.0);
    let mut jd3 = B2prismaticJointDef::default();
    jd3.initialize(g.clone(), b3.clone(), B2vec2::new(2.5, 12.0), B2vec2::new(0.0, 1.0));
    jd3.lower_translation = -5.0;
    jd3.upper_translation = 5.0;
    jd3.enable_limit = true;
    let j3 = world.borrow_mut().create_joint(&B2JointDefEnum::PrismaticJoint(jd3));
    gear(world, &b1, &b2, &j1, &j2, 2.0);
    gear(world, &b2, &b3, &j2, &j3, -0.5);
    b1.borrow_mut().set_angular_velocity(6.0);
}
fn pulley(world: &World, a: &Body, b: &Body, ga: (f32, f32), gb: (f32, f32), pa: (f32, f32), pb: (f32, f32), ratio: f32) {
    let mut jd = B2pulleyJointDef::default();
    jd.initialize(a.clone(), b.clone(), B2vec2::new(ga.0, ga.1), B2vec2::new(gb.0, gb.1), B2vec2::new(pa.0, pa.1),
                  B2vec2::new(pb.0, pb.1), ratio);
    world.borrow_mut().create_joint(&B2JointDefEnum::PulleyJoint(jd));
}
fn pulleys(world: &World) {
    // scenes.py::pulleys — the testbed's PulleyJoint scene, a block-and-tackle pair over a floor, a mouse drag
    let (y, l, a, b) = (16.0f32, 12.0f32, 1.0f32, 2.0f32);
    let g = body(world, false, 0.0, 0.0, 0.0);
    for x in [-10.0f32, 10.0f32] {
        B2body::create_fixture_by_shape(g.clone(), circle_at(2.0, x, y + b + l), 0.0);
    }
    edge(world, &g, -60.0, 0.0, 60.0, 0.0);
    let shape = boxed(a, b);
    let body1 = body(world, true, -10.0, y as f64, 0.0);
    B2body::create_fixture_by_shape(body1.clone(), shape.clone(), 5.0);
    let body2 = body(world, true, 10.0, y as f64, 0.0);
    B2body::create_fixture_by_shape(body2.clone(), shape, 5.0);
    pulley(world, &body1, &body2, (-10.0, y + b + l), (10.0, y + b + l), (-10.0, y + b), (10.0, y + b), 1.5);
    let (small, big) = (boxed(0.5, 0.5), boxed(1.0, 1.0));
    let body3 = body(world, true, 24.0, 6.0, 0.2);
    fixture(&body3, small, 1.0, 0.4);
    let body4 = body(world, true, 32.0, 9.0, 0.0);
    fixture(&body4, big.clone(), 2.0, 0.4);
    pulley(world, &body3, &body4, (25.0, 20.0), (31.0, 20.0), (24.0, 6.5), (32.0, 10.0), 2.0);
    let crate_ = body(world, true, -30.0, 1.0, 0.0);
    fixture(&crate_, big, 1.0, 0.5);
    let mut jd = B2mouseJointDef::default();
    jd.base.body_a = Some(g.clone());
    jd.base.body_b = Some(crate_.clone());
    jd.target.set(-29.5, 1.5);
    jd.max_force = 4000.0; // 1000 * mass
    b2_linear_stiffness(&mut jd.stiffness, &mut jd.damping, 5.0, 0.7, g.clone(), crate_.clone());
    let mouse = world.borrow_mut().create_joint(&B2JointDefEnum::MouseJoint(jd));
    if let JointAsDerivedMut::EMouseJoint(m) = mouse.borrow_mut().as_derived_mut() {
        m.set_target(B2vec2::new(-22.0, 9.0));
    }
    crate_.borrow_mut().set_awake(true);
}

fn main() {
    let out = PathBuf::from(std::env::args().nth(1).unwrap_or_else(|| "reference_dump".to_string()));
    std::fs::create_dir_all(&out).unwrap();
    // name, gravity, recipe, steps at which the state is written (0 = as built) — keep in sync with
    // tests/test_reference_dump.py::CASES
    let cases: Vec<(&str, (f32, f32), Box<dyn Fn(&World)>, Vec<usize>)> = vec![
        ("hello_world", (0.0, -10.0), Box::new(hello_world), vec![0, 1, 30, 60, 90]),
        ("pyramid", (0.0, -10.0), Box::new(pyramid), vec![0, 1, 2, 10, 60, 200, 300, 1000]),
        ("pile400", (0.0, -10.0), Box::new(|w| pile(w, 400, 12.0)), vec![0, 1, 50, 150, 200]),
        ("addpair2000", (0.0, 0.0), Box::new(|w| add_pair(w, 2000)), vec![0, 1, 40, 120, 150]),
        ("bridge", (0.0, -10.0), Box::new(bridge), vec![0, 1, 60, 240]),
        ("tumbler", (0.0, -10.0), Box::new(|w| tumbler(w, 120)), vec![0, 1, 60, 240]),
        ("pendulum", (0.0, -10.0), Box::new(pendulum), vec![0, 1, 60, 240]),
        ("gears", (0.0, -10.0), Box::new(gears), vec![0, 1, 30, 120, 300]),
        ("pulleys", (0.0, -10.0), Box::new(pulleys), vec![0, 1, 60, 150, 300]),
    ];
    let dt: f32 = 1.0 / 60.0;
    for (name, g, recipe, steps) in cases {
        let world = B2world::<Ud>::new(B2vec2::new(g.0, g.1));
        world.borrow_mut().set_continuous_physics(false); // TOI is out of scope in both engines
        recipe(&world);
        let mut done = 0usize;
        for s in steps {
            while done < s {
                world.borrow_mut().step(dt, 8, 3);
                done += 1;
            }
            let snap = world.borrow().gpu_snapshot(); // = private::gpu::mirror::flatten(&world)
            snap.save(&out.join(format!("{}_{:04}.b2snap", name, s))).unwrap();
        }
        println!("{}: {} steps dumped", name, done);
    }
}
