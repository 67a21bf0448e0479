// TEST INFRASTRUCTURE — CPU oracle (see b2o_math.hpp header). PARITY UNPINNED beyond the
// reference's own tests.
//
// b2o_tree.hpp — restates src/b2_dynamic_tree.rs + src/private/collision/b2_dynamic_tree.rs
// and src/b2_broad_phase.rs + src/private/collision/b2_broad_phase.rs.
#pragma once
#include <utility>
#include <vector>

#include "b2o_math.hpp"

namespace b2o {

constexpr int NULL_NODE = -1;

struct TreeNode {  // src/b2_dynamic_tree.rs:11-32
  AABB aabb;
  int user_data = -1;  // proxy index
  int parent = NULL_NODE;  // or next free
  int child1 = NULL_NODE, child2 = NULL_NODE;
  int height = -1;
  bool moved = false;
  bool is_leaf() const { return child1 == NULL_NODE; }
};

struct DynamicTree {
  int root = NULL_NODE;
  std::vector<TreeNode> nodes;
  int node_count = 0, node_capacity = 16, free_list = 0, insertion_count = 0;

  DynamicTree() {  // private :7-32
    nodes.resize(node_capacity);
    for (int i = 0; i < node_capacity - 1; ++i) { nodes[i].parent = i + 1; nodes[i].height = -1; }
    nodes[node_capacity - 1].parent = NULL_NODE;
    nodes[node_capacity - 1].height = -1;
  }
  int allocate_node() {  // :34-66
    if (free_list == NULL_NODE) {
      node_capacity *= 2;
      nodes.resize(node_capacity);
      for (int i = node_count; i < node_capacity - 1; ++i) { nodes[i].parent = i + 1; nodes[i].height = -1; }
      nodes[node_capacity - 1].parent = NULL_NODE;
      nodes[node_capacity - 1].height = -1;
      free_list = node_count;
    }
    int id = free_list;
    free_list = nodes[id].parent;
    nodes[id].parent = NULL_NODE;
    nodes[id].child1 = NULL_NODE;
    nodes[id].child2 = NULL_NODE;
    nodes[id].height = 0;
    nodes[id].user_data = -1;
    nodes[id].moved = false;
    ++node_count;
    return id;
  }
  void free_node(int id) {  // :69-76
    nodes[id].parent = free_list;
    nodes[id].height = -1;
    free_list = id;
    --node_count;
  }
  int create_proxy(const AABB& aabb, int user_data) {  // :81-99
    int id = allocate_node();
    Vec2 r(AABB_EXTENSION, AABB_EXTENSION);
    nodes[id].aabb.lower = aabb.lower - r;
    nodes[id].aabb.upper = aabb.upper + r;
    nodes[id].user_data = user_data;
    nodes[id].height = 0;
    nodes[id].moved = true;
    insert_leaf(id);
    return id;
  }
  void destroy_proxy(int id) { remove_leaf(id); free_node(id); }
  bool move_proxy(int id, const AABB& aabb, Vec2 displacement) {  // :109-168
    Vec2 r(AABB_EXTENSION, AABB_EXTENSION);
    AABB fat;
    fat.lower = aabb.lower - r;
    fat.upper = aabb.upper + r;
    Vec2 d = AABB_MULTIPLIER * displacement;
    if (d.x < 0.0f) fat.lower.x += d.x; else fat.upper.x += d.x;
    if (d.y < 0.0f) fat.lower.y += d.y; else fat.upper.y += d.y;
    const AABB tree_aabb = nodes[id].aabb;
    if (tree_aabb.contains(aabb)) {
      AABB huge;
      huge.lower = fat.lower - 4.0f * r;
      huge.upper = fat.upper + 4.0f * r;
      if (huge.contains(tree_aabb)) return false;
    }
    remove_leaf(id);
    nodes[id].aabb = fat;
    insert_leaf(id);
    nodes[id].moved = true;
    return true;
  }
  void insert_leaf(int leaf) {  // :170-298
    ++insertion_count;
    if (root == NULL_NODE) { root = leaf; nodes[root].parent = NULL_NODE; return; }
    AABB leaf_aabb = nodes[leaf].aabb;
    int index = root;
    while (!nodes[index].is_leaf()) {
      int child1 = nodes[index].child1, child2 = nodes[index].child2;
      float area = nodes[index].aabb.get_perimeter();
      AABB combined;
      combined.combine_two(nodes[index].aabb, leaf_aabb);
      float combined_area = combined.get_perimeter();
      float cost = 2.0f * combined_area;
      float inheritance_cost = 2.0f * (combined_area - area);
      float cost1;
      if (nodes[child1].is_leaf()) {
        AABB aabb;
        aabb.combine_two(leaf_aabb, nodes[child1].aabb);
        cost1 = aabb.get_perimeter() + inheritance_cost;
      } else {
        AABB aabb;
        aabb.combine_two(leaf_aabb, nodes[child1].aabb);
        float old_area = nodes[child1].aabb.get_perimeter();
        float new_area = aabb.get_perimeter();
        cost1 = (new_area - old_area) + inheritance_cost;
      }
      float cost2;
      if (nodes[child2].is_leaf()) {
        AABB aabb;
        aabb.combine_two(leaf_aabb, nodes[child2].aabb);
        cost2 = aabb.get_perimeter() + inheritance_cost;
      } else {
        AABB aabb;
        aabb.combine_two(leaf_aabb, nodes[child2].aabb);
        float old_area = nodes[child2].aabb.get_perimeter();
        float new_area = aabb.get_perimeter();
        cost2 = new_area - old_area + inheritance_cost;
      }
      if (cost < cost1 && cost < cost2) break;
      index = cost1 < cost2 ? child1 : child2;
    }
    int sibling = index;
    int old_parent = nodes[sibling].parent;
    int new_parent = allocate_node();
    nodes[new_parent].parent = old_parent;
    nodes[new_parent].user_data = -1;
    nodes[new_parent].aabb.combine_two(leaf_aabb, nodes[sibling].aabb);
    nodes[new_parent].height = nodes[sibling].height + 1;
    if (old_parent != NULL_NODE) {
      if (nodes[old_parent].child1 == sibling) nodes[old_parent].child1 = new_parent;
      else nodes[old_parent].child2 = new_parent;
      nodes[new_parent].child1 = sibling;
      nodes[new_parent].child2 = leaf;
      nodes[sibling].parent = new_parent;
      nodes[leaf].parent = new_parent;
    } else {
      nodes[new_parent].child1 = sibling;
      nodes[new_parent].child2 = leaf;
      nodes[sibling].parent = new_parent;
      nodes[leaf].parent = new_parent;
      root = new_parent;
    }
    index = nodes[leaf].parent;
    while (index != NULL_NODE) {
      index = balance(index);
      int child1 = nodes[index].child1, child2 = nodes[index].child2;
      nodes[index].height = 1 + b2_max(nodes[child1].height, nodes[child2].height);
      nodes[index].aabb.combine_two(nodes[child1].aabb, nodes[child2].aabb);
      index = nodes[index].parent;
    }
  }
  void remove_leaf(int leaf) {  // :300-353
    if (leaf == root) { root = NULL_NODE; return; }
    int parent = nodes[leaf].parent;
    int grand_parent = nodes[parent].parent;
    int sibling = nodes[parent].child1 == leaf ? nodes[parent].child2 : nodes[parent].child1;
    if (grand_parent != NULL_NODE) {
      if (nodes[grand_parent].child1 == parent) nodes[grand_parent].child1 = sibling;
      else nodes[grand_parent].child2 = sibling;
      nodes[sibling].parent = grand_parent;
      free_node(parent);
      int index = grand_parent;
      while (index != NULL_NODE) {
        index = balance(index);
        int child1 = nodes[index].child1, child2 = nodes[index].child2;
        nodes[index].aabb.combine_two(nodes[child1].aabb, nodes[child2].aabb);
        nodes[index].height = 1 + b2_max(nodes[child1].height, nodes[child2].height);
        index = nodes[index].parent;
      }
    } else {
      root = sibling;
      nodes[sibling].parent = NULL_NODE;
      free_node(parent);
    }
  }
  int balance(int i_a) {  // :357-489
    TreeNode& a = nodes[i_a];
    if (a.is_leaf() || a.height < 2) return i_a;
    int i_b = a.child1, i_c = a.child2;
    TreeNode& b = nodes[i_b];
    TreeNode& c = nodes[i_c];
    int bal = c.height - b.height;
    if (bal > 1) {
      int i_f = c.child1, i_g = c.child2;
      TreeNode& f = nodes[i_f];
      TreeNode& g = nodes[i_g];
      c.child1 = i_a;
      c.parent = a.parent;
      a.parent = i_c;
      if (c.parent != NULL_NODE) {
        if (nodes[c.parent].child1 == i_a) nodes[c.parent].child1 = i_c;
        else nodes[c.parent].child2 = i_c;
      } else {
        root = i_c;
      }
      if (f.height > g.height) {
        c.child2 = i_f;
        a.child2 = i_g;
        g.parent = i_a;
        a.aabb.combine_two(b.aabb, g.aabb);
        c.aabb.combine_two(a.aabb, f.aabb);
        a.height = 1 + b2_max(b.height, g.height);
        c.height = 1 + b2_max(a.height, f.height);
      } else {
        c.child2 = i_g;
        a.child2 = i_f;
        f.parent = i_a;
        a.aabb.combine_two(b.aabb, f.aabb);
        c.aabb.combine_two(a.aabb, g.aabb);
        a.height = 1 + b2_max(b.height, f.height);
        c.height = 1 + b2_max(a.height, g.height);
      }
      return i_c;
    }
    if (bal < -1) {
      int i_d = b.child1, i_e = b.child2;
      TreeNode& d = nodes[i_d];
      TreeNode& e = nodes[i_e];
      b.child1 = i_a;
      b.parent = a.parent;
      a.parent = i_b;
      if (b.parent != NULL_NODE) {
        if (nodes[b.parent].child1 == i_a) nodes[b.parent].child1 = i_b;
        else nodes[b.parent].child2 = i_b;
      } else {
        root = i_b;
      }
      if (d.height > e.height) {
        b.child2 = i_d;
        a.child1 = i_e;
        e.parent = i_a;
        a.aabb.combine_two(c.aabb, e.aabb);
        b.aabb.combine_two(a.aabb, d.aabb);
        a.height = 1 + b2_max(c.height, e.height);
        b.height = 1 + b2_max(a.height, d.height);
      } else {
        b.child2 = i_e;
        a.child1 = i_d;
        d.parent = i_a;
        a.aabb.combine_two(c.aabb, d.aabb);
        b.aabb.combine_two(a.aabb, e.aabb);
        a.height = 1 + b2_max(c.height, d.height);
        b.height = 1 + b2_max(a.height, e.height);
      }
      return i_b;
    }
    return i_a;
  }
  // src/b2_dynamic_tree.rs:239-267 — explicit stack, pushes child1 then child2 (child2 popped first)
  template <class F> void query(F&& callback, const AABB& aabb) const {
    // B2growableStack<i32, 256>: inline storage, heap only beyond 256 entries (src/b2_growable_stack.rs)
    int inline_stack[256];
    std::vector<int> heap_stack;
    int* stack = inline_stack;
    int capacity = 256, count = 0;
    auto push = [&](int v) {
      if (count == capacity) {
        heap_stack.assign(stack, stack + count);
        heap_stack.resize((size_t)capacity * 2);
        stack = heap_stack.data();
        capacity *= 2;
      }
      stack[count++] = v;
    };
    push(root);
    while (count > 0) {
      int id = stack[--count];
      if (id == NULL_NODE) continue;
      const TreeNode& node = nodes[id];
      if (b2_test_overlap(node.aabb, aabb)) {
        if (node.is_leaf()) {
          if (!callback(id)) return;
        } else {
          push(node.child1);
          push(node.child2);
        }
      }
    }
  }
};

struct BroadPhase {  // src/b2_broad_phase.rs + private
  DynamicTree tree;
  int proxy_count = 0;
  std::vector<int> move_buffer;
  std::vector<std::pair<int, int>> pair_buffer;

  int create_proxy(const AABB& aabb, int user_data) {  // private :33-42
    int id = tree.create_proxy(aabb, user_data);
    ++proxy_count;
    move_buffer.push_back(id);
    return id;
  }
  void move_proxy(int id, const AABB& aabb, Vec2 displacement) {  // :50-59
    if (tree.move_proxy(id, aabb, displacement)) move_buffer.push_back(id);
  }
  void touch_proxy(int id) { move_buffer.push_back(id); }  // :62-64
  bool test_overlap(int a, int b) const { return b2_test_overlap(tree.nodes[a].aabb, tree.nodes[b].aabb); }

  // src/b2_broad_phase.rs:200-249 ; callback private :86-111
  template <class AddPair> void update_pairs(AddPair&& add_pair) {
    pair_buffer.clear();
    for (int q : move_buffer) {
      if (q == NULL_NODE) continue;
      const AABB fat = tree.nodes[q].aabb;
      tree.query(
          [&](int p) -> bool {
            if (p == q) return true;
            bool moved = tree.nodes[p].moved;
            if (moved && p > q) return true;
            pair_buffer.emplace_back(b2_min(p, q), b2_max(p, q));
            return true;
          },
          fat);
    }
    for (auto& pr : pair_buffer) add_pair(tree.nodes[pr.first].user_data, tree.nodes[pr.second].user_data);
    for (int q : move_buffer) {
      if (q == NULL_NODE) continue;
      tree.nodes[q].moved = false;
    }
    move_buffer.clear();
  }
};

}  // namespace b2o
