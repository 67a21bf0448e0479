// b2g_tree.h — per-world replica of the reference's incremental AABB tree.
//
// Reference: box2d-rs src/private/collision/b2_dynamic_tree.rs (:34 allocate_node, :69 free_node,
// :170 insert_leaf, :300 remove_leaf, :357 balance) and src/b2_dynamic_tree.rs:239-267 (query).
// Contact creation order in the reference is the order in which tree queries report leaves
// (src/b2_broad_phase.rs:200-249), which depends on the tree's topology; keeping an exact replica
// (same node ids, same free-list order, same rotations) is what makes the pair buffer come out in
// the reference's order.  One thread owns one world's tree; nodes live in strided global arrays.
#pragma once
#include "b2g_common.h"

namespace b2g {

struct Tree {
  float4* aabb;  // base pointers already offset to this world's element 0
  int4* link;    // parent child1 child2 height
  int* moved;
  int stride;
  int* ws;       // this world's scalar slot 0
  int ws_stride;
  int phys_cap;
  int* status;

  B2G_HD int& root() { return ws[WS_TREE_ROOT * ws_stride]; }
  B2G_HD int& free_list() { return ws[WS_TREE_FREE * ws_stride]; }
  B2G_HD int& count() { return ws[WS_TREE_COUNT * ws_stride]; }
  B2G_HD int& cap() { return ws[WS_TREE_CAP * ws_stride]; }
  B2G_HD int& insertions() { return ws[WS_TREE_INSERTIONS * ws_stride]; }

  B2G_HD int4 L(int i) const { return link[i * stride]; }
  B2G_HD void setL(int i, int4 v) { link[i * stride] = v; }
  B2G_HD Box A(int i) const { float4 a = aabb[i * stride]; Box b; b.lo = v2(a.x, a.y); b.hi = v2(a.z, a.w); return b; }
  B2G_HD void setA(int i, const Box& b) { aabb[i * stride] = make_float4(b.lo.x, b.lo.y, b.hi.x, b.hi.y); }
  B2G_HD int parent(int i) const { return link[i * stride].x; }
  B2G_HD int height(int i) const { return link[i * stride].w; }
  B2G_HD void set_parent(int i, int p) { link[i * stride].x = p; }
  B2G_HD void set_child1(int i, int c) { link[i * stride].y = c; }
  B2G_HD void set_child2(int i, int c) { link[i * stride].z = c; }
  B2G_HD void set_height(int i, int h) { link[i * stride].w = h; }

  B2G_HDN int allocate_node() {
    if (free_list() == -1) {
      // pool doubling (:36-52) inside the physical capacity
      int old_cap = cap();
      int new_cap = old_cap * 2;
      if (new_cap > phys_cap) { *status = B2GPU_E_CAPACITY; return -1; }
      int nc = count();
      for (int i = nc; i < new_cap - 1; ++i) setL(i, make_int4(i + 1, -1, -1, -1));
      setL(new_cap - 1, make_int4(-1, -1, -1, -1));
      cap() = new_cap;
      free_list() = nc;
    }
    int id = free_list();
    free_list() = parent(id);
    setL(id, make_int4(-1, -1, -1, 0));
    moved[id * stride] = 0;
    count() += 1;
    return id;
  }
  B2G_HDN void free_node(int id) {
    int4 l = L(id);
    l.x = free_list();
    l.w = -1;
    setL(id, l);
    free_list() = id;
    count() -= 1;
  }

  B2G_HDN int balance(int ia) {
    int4 a = L(ia);
    if (a.y == -1 || a.w < 2) return ia;
    int ib = a.y, ic = a.z;
    int4 b = L(ib), c = L(ic);
    int bal = c.w - b.w;
    if (bal > 1) {  // rotate C up
      int i_f = c.y, i_g = c.z;
      int4 f = L(i_f), g = L(i_g);
      c.y = ia;
      c.x = a.x;
      a.x = ic;
      if (c.x != -1) {
        if (L(c.x).y == ia) set_child1(c.x, ic); else set_child2(c.x, ic);
      } else {
        root() = ic;
      }
      Box ab = A(ib);
      if (f.w > g.w) {
        c.z = i_f;
        a.z = i_g;
        set_parent(i_g, ia);
        Box aa = box_union(ab, A(i_g));
        setA(ia, aa);
        setA(ic, box_union(aa, A(i_f)));
        a.w = 1 + imax(b.w, g.w);
        c.w = 1 + imax(a.w, f.w);
      } else {
        c.z = i_g;
        a.z = i_f;
        set_parent(i_f, ia);
        Box aa = box_union(ab, A(i_f));
        setA(ia, aa);
        setA(ic, box_union(aa, A(i_g)));
        a.w = 1 + imax(b.w, f.w);
        c.w = 1 + imax(a.w, g.w);
      }
      setL(ia, a);
      setL(ic, c);
      return ic;
    }
    if (bal < -1) {  // rotate B up
      int i_d = b.y, i_e = b.z;
      int4 d = L(i_d), e = L(i_e);
      b.y = ia;
      b.x = a.x;
      a.x = ib;
      if (b.x != -1) {
        if (L(b.x).y == ia) set_child1(b.x, ib); else set_child2(b.x, ib);
      } else {
        root() = ib;
      }
      Box ac = A(ic);
      if (d.w > e.w) {
        b.z = i_d;
        a.y = i_e;
        set_parent(i_e, ia);
        Box aa = box_union(ac, A(i_e));
        setA(ia, aa);
        setA(ib, box_union(aa, A(i_d)));
        a.w = 1 + imax(c.w, e.w);
        b.w = 1 + imax(a.w, d.w);
      } else {
        b.z = i_e;
        a.y = i_d;
        set_parent(i_d, ia);
        Box aa = box_union(ac, A(i_d));
        setA(ia, aa);
        setA(ib, box_union(aa, A(i_e)));
        a.w = 1 + imax(c.w, d.w);
        b.w = 1 + imax(a.w, e.w);
      }
      setL(ia, a);
      setL(ib, b);
      return ib;
    }
    return ia;
  }

  // The walk from a changed node to the root (insert_leaf :261-288, remove_leaf :333-346): balance, then height and box from
  // the two children.  Every node on it is internal.  balance() is entered only when the heights it would read ask for a
  // rotation; otherwise the node, its children's links and boxes and — early — the parent's link are loaded in one round
  // (the rotation-free level changes nothing the parent's link holds), so a level costs one dependent load instead of four.
  // The values stored are those of the plain loop.
  B2G_HD void refit_up(int index) {
    if (index == -1) return;
    int4 a = L(index);
    for (;;) {
      int4 b = L(a.y), c = L(a.z);
      Box ab = A(a.y), ac = A(a.z);
      int4 pa = a;
      if (a.x != -1) pa = L(a.x);
      if (a.w >= 2 && (c.w - b.w > 1 || c.w - b.w < -1)) {
        index = balance(index);
        a = L(index);
        b = L(a.y); c = L(a.z);
        ab = A(a.y); ac = A(a.z);
        if (a.x != -1) pa = L(a.x);
      }
      a.w = 1 + imax(b.w, c.w);
      setL(index, a);
      setA(index, box_union(ab, ac));
      if (a.x == -1) break;
      index = a.x;
      a = pa;
    }
  }

  B2G_HDN void insert_leaf(int leaf) {
    insertions() += 1;
    if (root() == -1) {
      root() = leaf;
      set_parent(leaf, -1);
      return;
    }
    Box leaf_box = A(leaf);
    // The descent keeps the chosen child's link and box (loaded to price it) for the next level: one dependent load round per
    // level instead of three.  Same arithmetic, same comparisons.
    int index = root();
    int4 n = L(index);
    Box nb = A(index);
    for (;;) {
      if (n.y == -1) break;
      const int child1 = n.y, child2 = n.z;
      const Box cb1 = A(child1), cb2 = A(child2);
      const int4 l1 = L(child1), l2 = L(child2);
      float area = box_perimeter(nb);
      float combined_area = box_perimeter(box_union(nb, leaf_box));
      float cost = 2.0f * combined_area;
      float inheritance = 2.0f * (combined_area - area);
      float cost1, cost2;
      {
        float na = box_perimeter(box_union(leaf_box, cb1));
        if (l1.y == -1) cost1 = na + inheritance;
        else cost1 = (na - box_perimeter(cb1)) + inheritance;
      }
      {
        float na = box_perimeter(box_union(leaf_box, cb2));
        if (l2.y == -1) cost2 = na + inheritance;
        else cost2 = na - box_perimeter(cb2) + inheritance;
      }
      if (cost < cost1 && cost < cost2) break;
      if (cost1 < cost2) { index = child1; n = l1; nb = cb1; }
      else { index = child2; n = l2; nb = cb2; }
    }
    int sibling = index;
    int old_parent = parent(sibling);
    int new_parent = allocate_node();
    if (new_parent < 0) return;
    setA(new_parent, box_union(leaf_box, A(sibling)));
    setL(new_parent, make_int4(old_parent, sibling, leaf, height(sibling) + 1));
    if (old_parent != -1) {
      if (L(old_parent).y == sibling) set_child1(old_parent, new_parent); else set_child2(old_parent, new_parent);
    } else {
      root() = new_parent;
    }
    set_parent(sibling, new_parent);
    set_parent(leaf, new_parent);
    refit_up(new_parent);
  }

  B2G_HDN void remove_leaf(int leaf) {
    if (leaf == root()) { root() = -1; return; }
    int par = parent(leaf);
    int4 pl = L(par);
    int grand = pl.x;
    int sibling = pl.y == leaf ? pl.z : pl.y;
    if (grand != -1) {
      if (L(grand).y == par) set_child1(grand, sibling); else set_child2(grand, sibling);
      set_parent(sibling, grand);
      free_node(par);
      refit_up(grand);
    } else {
      root() = sibling;
      set_parent(sibling, -1);
      free_node(par);
    }
  }

  // B2dynamicTree::create_proxy (:81-99): caller supplies the already-fattened box.
  B2G_HDN int create_leaf(const Box& fat) {
    int id = allocate_node();
    if (id < 0) return id;
    setA(id, fat);
    moved[id * stride] = 1;
    insert_leaf(id);
    return id;
  }
};

// Explicit-stack query (src/b2_dynamic_tree.rs:239-267): pushes child1 then child2, so child2 is
// visited first.  The reference's stack grows on demand (B2growableStack<i32,256>); kStack bounds it here.
#define B2G_QUERY_STACK 128

}  // namespace b2g
