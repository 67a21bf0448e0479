// b2g_checkpoint.cu — on-disk form of a b2gpu_snapshot (second half): checkpoint / resume of
// the full step state.  The reference's serde support (src/serialize/serialize_b2_world.rs:133-178) writes the
// world *definition* — gravity, body defs, fixture defs, joints — and a world rebuilt from it starts with an empty
// contact list, a freshly balanced tree and zero warm-start impulses, so its trajectory differs from the run that
// was saved.  A snapshot carries what the step actually reads (contacts in creation order with manifolds and
// impulses, the tree pool with its free list, the move buffer, sleep timers, m_inv_dt0), so a resumed run is
// bit-identical to an uninterrupted one.
//
// Host-only code (stdio + memcpy): it needs no device and is the same in the product library and in the test
// simulator.  File layout, little-endian, version 1:
//   FileHeader (fixed size, carries its own checksum) | bodies | fixtures | shapes | proxies | nodes | contacts |
//   move_buffer — the seven tables of b2gpu_snapshot, records exactly as declared in include/b2gpu.h.
// A reader rejects a file whose magic, version, endianness tag, record sizes or either checksum do not match, or
// whose indices point outside their tables (b2gpu_snapshot_validate), before anything reaches the device.
#include <fcntl.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <new>
#include <string>
#include <vector>

#include "b2g_runtime.h"

using namespace b2g;

namespace {

const char kMagic[8] = {'B', '2', 'G', 'P', 'U', 'S', 'N', 'P'};
const uint32_t kFileVersion = 2;  // 2: joint table after the move buffer (a version-1 file is a version-2 file without joints)
const uint32_t kEndianTag = 0x01020304u;

struct FileHeader {
  char magic[8];
  uint32_t file_version;
  uint32_t abi_version;
  uint32_t endian_tag;
  uint32_t header_bytes;
  uint32_t record_bytes[8];  // body, fixture, shape, proxy, node, contact, move-buffer entry, joint (0 in version-1 files)
  b2gpu_world_rec world;
  b2gpu_snapshot_sizes n;
  uint64_t payload_bytes;
  uint64_t payload_hash;  // FNV-1a 64 over the payload
  uint64_t header_hash;   // FNV-1a 64 over every header byte before this field
};

const int kTables = 8;
const uint32_t kRecordBytes[kTables] = {sizeof(b2gpu_body_rec),    sizeof(b2gpu_fixture_rec),   sizeof(b2gpu_shape_rec),
                                        sizeof(b2gpu_proxy_rec),   sizeof(b2gpu_tree_node_rec), sizeof(b2gpu_contact_rec),
                                        sizeof(int32_t),           sizeof(b2gpu_joint_rec)};

uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

struct Table { const void* ptr; size_t bytes; };

void tables_of(const b2gpu_snapshot* s, Table t[kTables]) {
  const b2gpu_snapshot_sizes& n = s->n;
  t[0] = {s->bodies, sizeof(b2gpu_body_rec) * (size_t)n.body_count};
  t[1] = {s->fixtures, sizeof(b2gpu_fixture_rec) * (size_t)n.fixture_count};
  t[2] = {s->shapes, sizeof(b2gpu_shape_rec) * (size_t)n.shape_count};
  t[3] = {s->proxies, sizeof(b2gpu_proxy_rec) * (size_t)n.proxy_count};
  t[4] = {s->nodes, sizeof(b2gpu_tree_node_rec) * (size_t)n.node_count};
  t[5] = {s->contacts, sizeof(b2gpu_contact_rec) * (size_t)n.contact_count};
  t[6] = {s->move_buffer, sizeof(int32_t) * (size_t)n.move_count};
  t[7] = {s->joints, sizeof(b2gpu_joint_rec) * (size_t)n.joint_count};
}

bool sizes_ok(const b2gpu_snapshot_sizes& n) {
  return n.body_count >= 0 && n.fixture_count >= 0 && n.shape_count >= 0 && n.proxy_count >= 0 && n.node_count >= 0 &&
         n.contact_count >= 0 && n.move_count >= 0 && n.joint_count >= 0;
}

int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
}

// Reads and checks the header; leaves the stream at the first payload byte.
int read_header(FILE* f, const char* path, FileHeader* h) {
  if (fread(h, 1, sizeof(FileHeader), f) != sizeof(FileHeader))
    return fail(B2GPU_E_INVALID, std::string("checkpoint: ") + path + " is shorter than a snapshot header");
  if (memcmp(h->magic, kMagic, 8) != 0) return fail(B2GPU_E_INVALID, std::string("checkpoint: ") + path + " is not a b2gpu snapshot file");
  if (h->endian_tag != kEndianTag) return fail(B2GPU_E_INVALID, "checkpoint: file was written with another byte order");
  if (h->file_version != kFileVersion && h->file_version != 1) return fail(B2GPU_E_UNSUPPORTED, "checkpoint: unknown file version " + std::to_string(h->file_version));
  if (h->header_bytes != sizeof(FileHeader)) return fail(B2GPU_E_INVALID, "checkpoint: header size mismatch");
  if (h->header_hash != fnv1a(h, offsetof(FileHeader, header_hash))) return fail(B2GPU_E_INVALID, "checkpoint: header checksum mismatch (corrupt file)");
  // ABI 2 added the joint table and changed no other record: version-1 files (ABI 1, no joints) still load
  if (h->abi_version != B2GPU_ABI_VERSION && !(h->abi_version == 1 && h->file_version == 1))
    return fail(B2GPU_E_UNSUPPORTED, "checkpoint: file carries ABI version " + std::to_string(h->abi_version) + ", this library is " +
                                         std::to_string(B2GPU_ABI_VERSION));
  const int n_tables = h->file_version == 1 ? 7 : kTables;
  for (int i = 0; i < n_tables; ++i)
    if (h->record_bytes[i] != kRecordBytes[i]) return fail(B2GPU_E_INVALID, "checkpoint: record layout differs from include/b2gpu.h");
  if (h->file_version == 1 && (h->record_bytes[7] != 0 || h->n.joint_count != 0)) return fail(B2GPU_E_INVALID, "checkpoint: version-1 file with a joint table");
  if (!sizes_ok(h->n)) return fail(B2GPU_E_INVALID, "checkpoint: negative table size");
  return 0;
}

#define GUARD_BEGIN try {
#define GUARD_END                                          \
  }                                                        \
  catch (const std::bad_alloc&) {                          \
    return fail(B2GPU_E_INVALID, "out of host memory");    \
  }                                                        \
  catch (...) {                                            \
    return fail(B2GPU_E_INVALID, "unexpected C++ exception"); \
  }

#define CHECK_RANGE(cond, what, i)                                                                                  \
  if (!(cond)) return fail(B2GPU_E_INVALID, std::string("snapshot: ") + what + " out of range at record " + std::to_string(i));

}  // namespace

extern "C" {

// Consistency of a snapshot: every link stays inside its table AND the linked structures are what the step assumes
// (acyclic fixture lists owned by one body each, a tree whose internal nodes have two children pointing back and whose
// leaves own a proxy, a free list covering the rest of the pool, move-buffer entries that are leaves).  Not a physics
// check — a snapshot that passes cannot make upload or a step loop forever or read outside the arrays.
int b2gpu_snapshot_validate(const b2gpu_snapshot* s) {
  if (!s) return fail(B2GPU_E_INVALID, "snapshot_validate: snapshot is NULL");
  const b2gpu_snapshot_sizes& n = s->n;
  if (!sizes_ok(n)) return fail(B2GPU_E_INVALID, "snapshot: negative table size");
  Table t[kTables];
  tables_of(s, t);
  for (int i = 0; i < kTables; ++i)
    if (t[i].bytes && !t[i].ptr) return fail(B2GPU_E_INVALID, "snapshot: a non-empty table has a NULL pointer");
  const int nb = n.body_count, nf = n.fixture_count, ns = n.shape_count, np = n.proxy_count, nn = n.node_count;
  for (int i = 0; i < nb; ++i) {
    const b2gpu_body_rec& b = s->bodies[i];
    CHECK_RANGE(b.type >= B2GPU_STATIC_BODY && b.type <= B2GPU_DYNAMIC_BODY, "body type", i);
    CHECK_RANGE(b.fixture_head >= -1 && b.fixture_head < nf, "body fixture_head", i);
    CHECK_RANGE(b.fixture_count >= 0 && b.fixture_count <= nf, "body fixture_count", i);
  }
  for (int i = 0; i < nf; ++i) {
    const b2gpu_fixture_rec& f = s->fixtures[i];
    CHECK_RANGE(f.body >= 0 && f.body < nb, "fixture body", i);
    CHECK_RANGE(f.next >= -1 && f.next < nf, "fixture next", i);
    CHECK_RANGE(f.shape_type >= B2GPU_SHAPE_CIRCLE && f.shape_type <= B2GPU_SHAPE_CHAIN, "fixture shape_type", i);
    CHECK_RANGE(f.child_count >= 1 && f.shape_first >= 0 && (long long)f.shape_first + f.child_count <= ns, "fixture shape range", i);
    CHECK_RANGE(f.proxy_first >= -1 && (f.proxy_first < 0 || (long long)f.proxy_first + f.child_count <= np), "fixture proxy range", i);
  }
  for (int i = 0; i < ns; ++i) {
    const b2gpu_shape_rec& sh = s->shapes[i];
    CHECK_RANGE(sh.type >= B2GPU_SHAPE_CIRCLE && sh.type <= B2GPU_SHAPE_POLYGON, "shape type", i);
    CHECK_RANGE(sh.type != B2GPU_SHAPE_POLYGON || (sh.count >= 1 && sh.count <= B2GPU_MAX_POLYGON_VERTICES), "polygon vertex count", i);
  }
  for (int i = 0; i < np; ++i) {
    const b2gpu_proxy_rec& p = s->proxies[i];
    CHECK_RANGE(p.fixture >= 0 && p.fixture < nf, "proxy fixture", i);
    CHECK_RANGE(p.child_index >= 0 && p.child_index < s->fixtures[p.fixture].child_count, "proxy child_index", i);
    CHECK_RANGE(p.proxy_id >= -1 && p.proxy_id < nn, "proxy tree node", i);
  }
  for (int i = 0; i < nn; ++i) {
    const b2gpu_tree_node_rec& d = s->nodes[i];
    CHECK_RANGE(d.parent >= -1 && d.parent < nn, "tree node parent", i);
    CHECK_RANGE(d.child1 >= -1 && d.child1 < nn && d.child2 >= -1 && d.child2 < nn, "tree node child", i);
    CHECK_RANGE(d.proxy >= -1 && d.proxy < np, "tree node proxy", i);
    CHECK_RANGE(d.height >= -1, "tree node height", i);
  }
  for (int i = 0; i < n.contact_count; ++i) {
    const b2gpu_contact_rec& c = s->contacts[i];
    CHECK_RANGE(c.fixture_a >= 0 && c.fixture_a < nf && c.fixture_b >= 0 && c.fixture_b < nf, "contact fixture", i);
    CHECK_RANGE(c.index_a >= 0 && c.index_a < s->fixtures[c.fixture_a].child_count, "contact index_a", i);
    CHECK_RANGE(c.index_b >= 0 && c.index_b < s->fixtures[c.fixture_b].child_count, "contact index_b", i);
    CHECK_RANGE(c.manifold.point_count >= 0 && c.manifold.point_count <= 2, "manifold point_count", i);
    CHECK_RANGE(c.manifold.type >= B2GPU_MANIFOLD_CIRCLES && c.manifold.type <= B2GPU_MANIFOLD_FACE_B, "manifold type", i);
  }
  for (int i = 0; i < n.move_count; ++i) CHECK_RANGE(s->move_buffer[i] >= -1 && s->move_buffer[i] < nn, "move buffer entry", i);
  for (int i = 0; i < n.joint_count; ++i) {
    const b2gpu_joint_rec& j = s->joints[i];
    CHECK_RANGE(j.type == B2GPU_JOINT_REVOLUTE || j.type == B2GPU_JOINT_DISTANCE || j.type == B2GPU_JOINT_WELD || j.type == B2GPU_JOINT_PRISMATIC || j.type == B2GPU_JOINT_WHEEL || j.type == B2GPU_JOINT_FRICTION || j.type == B2GPU_JOINT_MOTOR || j.type == B2GPU_JOINT_PULLEY || j.type == B2GPU_JOINT_MOUSE || j.type == B2GPU_JOINT_GEAR, "joint type", i);
    CHECK_RANGE(j.body_a >= 0 && j.body_a < nb && j.body_b >= 0 && j.body_b < nb && j.body_a != j.body_b, "joint body", i);
    if (j.type == B2GPU_JOINT_GEAR) {
      int32_t bc, bd;
      memcpy(&bc, &j.impulse[5], 4); memcpy(&bd, &j.impulse[6], 4);
      CHECK_RANGE(bc >= 0 && bc < nb && bd >= 0 && bd < nb, "gear joint body C / D", i);
    }
  }
  const b2gpu_world_rec& w = s->world;
  if (w.tree_root < -1 || w.tree_root >= nn || w.tree_free_list < -1 || w.tree_free_list >= nn || w.tree_node_capacity != nn ||
      w.tree_node_count < 0 || w.tree_node_count > nn || w.proxy_count < 0 || w.proxy_count > np)
    return fail(B2GPU_E_INVALID, "snapshot: world tree bookkeeping inconsistent with the node table");
  GUARD_BEGIN
  // ---- structure (range checks alone let cycles and dangling links through: a fixture list that loops makes the
  //      host-side builders spin, a half-linked tree node makes the device read link[-1]).
  // (1) every fixture sits on exactly one body's list, the list of its own body, and the counts agree
  std::vector<unsigned char> seen_f((size_t)nf, 0);
  for (int b = 0; b < nb; ++b) {
    int count = 0;
    for (int f = s->bodies[b].fixture_head; f != -1; f = s->fixtures[f].next) {
      if (seen_f[f]) return fail(B2GPU_E_INVALID, "snapshot: fixture lists form a cycle or share fixture " + std::to_string(f));
      seen_f[f] = 1;
      if (s->fixtures[f].body != b) return fail(B2GPU_E_INVALID, "snapshot: fixture " + std::to_string(f) + " is on the list of another body");
      ++count;
    }
    if (count != s->bodies[b].fixture_count) return fail(B2GPU_E_INVALID, "snapshot: fixture_count of body " + std::to_string(b) + " does not match its list");
  }
  for (int f = 0; f < nf; ++f)
    if (!seen_f[f]) return fail(B2GPU_E_INVALID, "snapshot: fixture " + std::to_string(f) + " is on no body's list");
  // (2) a fixture's proxies are its own, one per child, in child order
  for (int f = 0; f < nf; ++f) {
    const b2gpu_fixture_rec& fx = s->fixtures[f];
    for (int c = 0; fx.proxy_first >= 0 && c < fx.child_count; ++c) {
      const b2gpu_proxy_rec& p = s->proxies[fx.proxy_first + c];
      if (p.fixture != f || p.child_index != c) return fail(B2GPU_E_INVALID, "snapshot: proxy table does not match fixture " + std::to_string(f));
    }
  }
  for (int i = 0; i < n.contact_count; ++i)
    if (s->fixtures[s->contacts[i].fixture_a].proxy_first < 0 || s->fixtures[s->contacts[i].fixture_b].proxy_first < 0)
      return fail(B2GPU_E_INVALID, "snapshot: contact " + std::to_string(i) + " references a fixture without proxies");
  // (3) the tree: walked from the root, every node once; internal nodes have two children that point back, leaves
  //     carry a proxy that points back; as many nodes as the world record says
  std::vector<unsigned char> state((size_t)nn, 0);  // 1 = internal node of the tree, 2 = leaf, 3 = on the free list
  int in_tree = 0;
  if (w.tree_root != -1) {
    if (s->nodes[w.tree_root].parent != -1) return fail(B2GPU_E_INVALID, "snapshot: tree root has a parent");
    std::vector<int> stack(1, w.tree_root);
    while (!stack.empty()) {
      const int i = stack.back();
      stack.pop_back();
      if (state[i]) return fail(B2GPU_E_INVALID, "snapshot: tree node " + std::to_string(i) + " is reachable twice (cycle)");
      const b2gpu_tree_node_rec& d = s->nodes[i];
      ++in_tree;
      if (d.height < 0) return fail(B2GPU_E_INVALID, "snapshot: free node " + std::to_string(i) + " is linked into the tree");
      if (d.child1 == -1) {
        state[i] = 2;
        if (d.child2 != -1 || d.height != 0) return fail(B2GPU_E_INVALID, "snapshot: tree leaf " + std::to_string(i) + " is half-linked");
        if (d.proxy < 0 || s->proxies[d.proxy].proxy_id != i) return fail(B2GPU_E_INVALID, "snapshot: tree leaf " + std::to_string(i) + " has no proxy pointing back at it");
      } else {
        state[i] = 1;
        if (d.child2 == -1 || d.child1 == d.child2 || d.height < 1) return fail(B2GPU_E_INVALID, "snapshot: internal tree node " + std::to_string(i) + " is half-linked");
        if (s->nodes[d.child1].parent != i || s->nodes[d.child2].parent != i)
          return fail(B2GPU_E_INVALID, "snapshot: a child of tree node " + std::to_string(i) + " does not point back");
        stack.push_back(d.child1);
        stack.push_back(d.child2);
      }
    }
  }
  if (in_tree != w.tree_node_count) return fail(B2GPU_E_INVALID, "snapshot: tree_node_count does not match the nodes reachable from the root");
  int on_free = 0;
  for (int i = w.tree_free_list; i != -1; i = s->nodes[i].parent) {
    if (state[i]) return fail(B2GPU_E_INVALID, "snapshot: free list revisits node " + std::to_string(i));
    state[i] = 3;
    ++on_free;
  }
  if (on_free != nn - w.tree_node_count) return fail(B2GPU_E_INVALID, "snapshot: free list length does not match the node pool");
  // (4) proxies and move-buffer entries name leaves of that tree
  for (int i = 0; i < np; ++i) {
    const int id = s->proxies[i].proxy_id;
    if (id >= 0 && (state[id] != 2 || s->nodes[id].proxy != i)) return fail(B2GPU_E_INVALID, "snapshot: proxy " + std::to_string(i) + " does not own its tree leaf");
  }
  for (int i = 0; i < n.move_count; ++i)
    if (s->move_buffer[i] != -1 && state[s->move_buffer[i]] != 2) return fail(B2GPU_E_INVALID, "snapshot: move buffer entry " + std::to_string(i) + " is not a tree leaf");
  return 0;
  GUARD_END
}

// Writes `s` to `path` (through `path`.tmp + rename, so a crash mid-write never leaves a truncated checkpoint under
// the final name).
int b2gpu_snapshot_save(const b2gpu_snapshot* s, const char* path) {
  GUARD_BEGIN
  if (!s || !path || !*path) return fail(B2GPU_E_INVALID, "snapshot_save: bad argument");
  int rc = b2gpu_snapshot_validate(s);
  if (rc) return rc;
  Table t[kTables];
  tables_of(s, t);
  FileHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, kMagic, 8);
  h.file_version = kFileVersion;
  h.abi_version = B2GPU_ABI_VERSION;
  h.endian_tag = kEndianTag;
  h.header_bytes = sizeof(FileHeader);
  memcpy(h.record_bytes, kRecordBytes, sizeof kRecordBytes);
  h.world = s->world;
  h.n = s->n;
  uint64_t hash = 1469598103934665603ull;
  for (int i = 0; i < kTables; ++i) {
    h.payload_bytes += t[i].bytes;
    hash = fnv1a(t[i].ptr, t[i].bytes, hash);
  }
  h.payload_hash = hash;
  h.header_hash = fnv1a(&h, offsetof(FileHeader, header_hash));
  // unique temporary name (two writers saving to the same path do not collide), data on stable storage before the
  // rename and the directory entry after it: a crash never leaves a truncated file under the final name
  std::string tmp = std::string(path) + ".XXXXXX.tmp";
  const int fd = mkstemps(&tmp[0], 4);
  if (fd < 0) return fail(B2GPU_E_IO, "snapshot_save: cannot create a temporary file next to " + std::string(path));
  FILE* f = fdopen(fd, "wb");
  if (!f) { close(fd); remove(tmp.c_str()); return fail(B2GPU_E_IO, "snapshot_save: cannot open " + tmp); }
  bool ok = fwrite(&h, 1, sizeof h, f) == sizeof h;
  for (int i = 0; ok && i < kTables; ++i) ok = t[i].bytes == 0 || fwrite(t[i].ptr, 1, t[i].bytes, f) == t[i].bytes;
  ok = (fflush(f) == 0) && ok;
  ok = (fsync(fileno(f)) == 0) && ok;
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {
    remove(tmp.c_str());
    return fail(B2GPU_E_IO, std::string("snapshot_save: write to ") + path + " failed");
  }
  {
    std::string dir(path);
    const size_t slash = dir.find_last_of('/');
    dir = slash == std::string::npos ? "." : (slash == 0 ? "/" : dir.substr(0, slash));
    const int dfd = open(dir.c_str(), O_RDONLY);
    if (dfd >= 0) { fsync(dfd); close(dfd); }
  }
  return 0;
  GUARD_END
}

// Table sizes (and world scalars) of a checkpoint, so the caller can allocate the arrays for b2gpu_snapshot_load.
int b2gpu_snapshot_file_sizes(const char* path, b2gpu_snapshot_sizes* out) {
  GUARD_BEGIN
  if (!path || !out) return fail(B2GPU_E_INVALID, "snapshot_file_sizes: bad argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(B2GPU_E_IO, std::string("snapshot_file_sizes: cannot open ") + path);
  FileHeader h;
  int rc = read_header(f, path, &h);
  fclose(f);
  if (rc) return rc;
  *out = h.n;
  return 0;
  GUARD_END
}

// Reads a checkpoint into caller-owned arrays; `out->n` holds their capacities on entry (as for
// b2gpu_world_download) and the table sizes on return.  Nothing is written to `out` unless the whole file checks out.
int b2gpu_snapshot_load(const char* path, b2gpu_snapshot* out) {
  GUARD_BEGIN
  if (!path || !out) return fail(B2GPU_E_INVALID, "snapshot_load: bad argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(B2GPU_E_IO, std::string("snapshot_load: cannot open ") + path);
  FileHeader h;
  int rc = read_header(f, path, &h);
  if (rc) { fclose(f); return rc; }
  const b2gpu_snapshot_sizes cap = out->n;
  if (h.n.body_count > cap.body_count || h.n.fixture_count > cap.fixture_count || h.n.shape_count > cap.shape_count ||
      h.n.proxy_count > cap.proxy_count || h.n.node_count > cap.node_count || h.n.contact_count > cap.contact_count ||
      h.n.move_count > cap.move_count || h.n.joint_count > cap.joint_count) {
    fclose(f);
    return fail(B2GPU_E_CAPACITY, "snapshot_load: caller arrays are smaller than the tables in the file (see b2gpu_snapshot_file_sizes)");
  }
  b2gpu_snapshot staged = *out;
  staged.world = h.world;
  staged.n = h.n;
  Table t[kTables];
  tables_of(&staged, t);
  uint64_t total = 0;
  for (int i = 0; i < kTables; ++i) {
    total += t[i].bytes;
    if (t[i].bytes && !t[i].ptr) { fclose(f); return fail(B2GPU_E_INVALID, "snapshot_load: a table of the file is not empty but the caller's array pointer is NULL"); }
  }
  if (total != h.payload_bytes) { fclose(f); return fail(B2GPU_E_INVALID, "checkpoint: payload size does not match the table sizes"); }
  std::vector<unsigned char> buf(total ? total : 1);
  size_t got = fread(buf.data(), 1, total, f);
  unsigned char extra;
  bool trailing = fread(&extra, 1, 1, f) == 1;
  fclose(f);
  if (got != total) return fail(B2GPU_E_INVALID, std::string("checkpoint: ") + path + " is truncated");
  if (trailing) return fail(B2GPU_E_INVALID, std::string("checkpoint: ") + path + " has bytes after the payload");
  if (fnv1a(buf.data(), total) != h.payload_hash) return fail(B2GPU_E_INVALID, "checkpoint: payload checksum mismatch (corrupt file)");
  // validate in the staging buffer, then copy out
  b2gpu_snapshot view = staged;
  size_t off = 0;
  unsigned char* base = buf.data();
  view.bodies = (b2gpu_body_rec*)(base + off); off += t[0].bytes;
  view.fixtures = (b2gpu_fixture_rec*)(base + off); off += t[1].bytes;
  view.shapes = (b2gpu_shape_rec*)(base + off); off += t[2].bytes;
  view.proxies = (b2gpu_proxy_rec*)(base + off); off += t[3].bytes;
  view.nodes = (b2gpu_tree_node_rec*)(base + off); off += t[4].bytes;
  view.contacts = (b2gpu_contact_rec*)(base + off); off += t[5].bytes;
  view.move_buffer = (int32_t*)(base + off); off += t[6].bytes;
  view.joints = (b2gpu_joint_rec*)(base + off);
  rc = b2gpu_snapshot_validate(&view);
  if (rc) return rc;
  off = 0;
  for (int i = 0; i < kTables; ++i) {
    if (t[i].bytes) memcpy(const_cast<void*>(t[i].ptr), base + off, t[i].bytes);
    off += t[i].bytes;
  }
  out->world = h.world;
  out->n = h.n;
  return 0;
  GUARD_END
}

}  // extern "C"
