// b2g_events.cu — B2contactListener::begin_contact / end_contact (src/b2_world_callbacks.rs:68-104) for a
// device-resident step ("buffered begin/end events").  The reference fires them inside the
// step: in `collide`, walking the contact list newest-first, B2contact::update fires begin_contact when a contact
// starts touching and end_contact when it stops (b2_contact.rs(private):201-211), and the contacts collected for
// destruction are destroyed after the loop, in the same order, each firing end_contact if it was touching
// (b2_contact_manager.rs(private):24-49, 164-170).  Nothing else in a step changes a touching flag, so the events
// of a step are a function of the contact tables before and after it; this file derives them on the host from the two
// snapshots, in the reference's firing order, for replay to a listener after the step.  Host-only code.
//
// The contact arrays are in creation order with stable compaction (include/b2gpu.h), so `after` is
// [survivors of `before`, in order] + [contacts created in this step, in creation order], and the collide loop's
// newest-first walk is: created by find_new_contacts at the top of the step (descending), then `before` (descending).
#include <stdint.h>

#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "b2g_runtime.h"

using namespace b2g;

namespace {

struct Key {
  uint64_t a, b;
  bool operator==(const Key& o) const { return a == o.a && b == o.b; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = k.a * 0x9E3779B97F4A7C15ull;
    h ^= (k.b + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};
Key key_of(const b2gpu_contact_rec& c) {
  return {((uint64_t)(uint32_t)c.fixture_a << 32) | (uint32_t)c.index_a, ((uint64_t)(uint32_t)c.fixture_b << 32) | (uint32_t)c.index_b};
}
int fail(int code, const char* msg) {
  set_error(msg);
  return code;
}

}  // namespace

extern "C" int b2gpu_contact_events(const b2gpu_snapshot* before, const b2gpu_snapshot* after, int destroyed,
                                    b2gpu_contact_event* out, int capacity) {
  try {
    if (!before || !after || capacity < 0 || (capacity > 0 && !out)) return fail(B2GPU_E_INVALID, "contact_events: bad argument");
    const int nb = before->n.contact_count, na = after->n.contact_count;
    if (nb < 0 || na < 0 || (nb > 0 && !before->contacts) || (na > 0 && !after->contacts))
      return fail(B2GPU_E_INVALID, "contact_events: bad contact table");
    // one contact per (fixture, child) pair: add_pair's duplicate test (b2_contact_manager.rs(private):196-224)
    std::unordered_map<Key, int, KeyHash> in_before;
    in_before.reserve((size_t)nb * 2 + 1);
    for (int i = 0; i < nb; ++i) in_before[key_of(before->contacts[i])] = i;
    // survivors = a prefix of `after`
    int n_surv;
    if (destroyed >= 0) {
      n_surv = nb - destroyed;
      if (n_surv < 0 || n_surv > na) return fail(B2GPU_E_INVALID, "contact_events: destroyed count inconsistent with the snapshots");
    } else {  // infer: longest prefix of `after` that is a subsequence of `before`
      int k = 0;
      n_surv = 0;
      while (n_surv < na && n_surv < nb) {  // survivors <= min(before, after); best effort: pass the step's `destroyed` stat
                                             // when a destroyed contact may have been re-created for the same fixture pair
        auto it = in_before.find(key_of(after->contacts[n_surv]));
        if (it == in_before.end() || it->second < k) break;
        k = it->second + 1;
        ++n_surv;
      }
    }
    int count = 0;
    auto emit = [&](int type, const b2gpu_contact_rec& c) {
      if (count < capacity) {
        b2gpu_contact_event& e = out[count];
        e.type = type;
        e.fixture_a = c.fixture_a; e.index_a = c.index_a;
        e.fixture_b = c.fixture_b; e.index_b = c.index_b;
        e.reserved[0] = e.reserved[1] = e.reserved[2] = 0;
      }
      ++count;
    };
    // collide loop, newest first: contacts created in this step (only those made by find_new_contacts at the top of
    // the step have been evaluated; the ones update_pairs made at its end are not touching yet)
    for (int j = na - 1; j >= n_surv; --j)
      if (after->contacts[j].flags & B2GPU_CONTACT_TOUCHING) emit(B2GPU_EVENT_BEGIN_CONTACT, after->contacts[j]);
    std::vector<char> survived((size_t)nb + 1, 0);
    std::vector<int> src((size_t)n_surv + 1, -1);
    for (int j = 0; j < n_surv; ++j) {
      auto it = in_before.find(key_of(after->contacts[j]));
      if (it == in_before.end() || survived[it->second]) return fail(B2GPU_E_INVALID, "contact_events: `after` is not a step of `before` (survivor not found)");
      survived[it->second] = 1;
      src[j] = it->second;
    }
    for (int j = n_surv - 1; j >= 0; --j) {
      const bool was = (before->contacts[src[j]].flags & B2GPU_CONTACT_TOUCHING) != 0;
      const bool now = (after->contacts[j].flags & B2GPU_CONTACT_TOUCHING) != 0;
      if (!was && now) emit(B2GPU_EVENT_BEGIN_CONTACT, after->contacts[j]);
      if (was && !now) emit(B2GPU_EVENT_END_CONTACT, after->contacts[j]);
    }
    // deferred destruction, in collection order (newest first)
    for (int i = nb - 1; i >= 0; --i)
      if (!survived[i] && (before->contacts[i].flags & B2GPU_CONTACT_TOUCHING)) emit(B2GPU_EVENT_END_CONTACT, before->contacts[i]);
    return count;
  } catch (const std::bad_alloc&) {
    return fail(B2GPU_E_INVALID, "out of host memory");
  } catch (...) {
    return fail(B2GPU_E_INVALID, "unexpected C++ exception");
  }
}
