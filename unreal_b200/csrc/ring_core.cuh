// Index logic of the replay ring, one env at a time (host+device so tests/cpu_harness can
// exercise it without a GPU).  Restates Experience in train/experience.py:48-153.
//
// Storage per env: H packed frame records in a circular buffer, slot(abs) = abs % H, where
// abs is the reference's absolute frame index (top + position in the deque).  The reference's
// two index deques (_pos_reward_indices / _neg_reward_indices, :52-55) always hold, in
// order, every absolute index in [max(3, top+3), top+count-1] split by reward sign (> 0 or
// not): an add appends the new index (:75-80) and, once full, advances top and pops the one
// index (top_old+3) that fell below the cut (:82-93).  So only their LENGTHS need storing;
// the k-th element is found by rank-select over the records.
#pragma once
#include "common.cuh"
#include "maze_core.cuh"

namespace unreal {

struct RingRef {
  uint64_t* rec;    // this env's H records
  int H;
  int64_t* top;     // _top_frame_index
  int32_t* count;   // len(_frames)
  int32_t* n_pos;   // len(_pos_reward_indices)
  int32_t* n_neg;   // len(_neg_reward_indices)
};

UNREAL_HD uint64_t ring_at_abs(const RingRef& r, int64_t abs_index) { return r.rec[abs_index % r.H]; }
UNREAL_HD uint64_t ring_at_raw(const RingRef& r, int raw) { return r.rec[(*r.top + raw) % r.H]; }

// add_frame (:63-93).  Returns 0 when the frame was discarded (invalid record, or a terminal
// directly after a terminal :64-67).
UNREAL_HD int ring_add(const RingRef& r, uint64_t frame) {
  if (!frame_valid(frame)) return 0;
  const int count = *r.count;
  const int64_t top = *r.top;
  if (frame_terminal(frame) && count > 0 && frame_terminal(ring_at_abs(r, top + count - 1))) return 0;
  const int64_t index = top + count;
  const bool was_full = count >= r.H;
  r.rec[index % r.H] = frame;
  if (index >= 3) {
    if (frame_reward(frame) > 0) *r.n_pos += 1; else *r.n_neg += 1;
  }
  if (was_full) {
    *r.top = top + 1;
    // cut = top_new + 3: exactly index top_old + 3 leaves the eligible range
    if (frame_reward(ring_at_abs(r, top + 3)) > 0) *r.n_pos -= 1; else *r.n_neg -= 1;
  } else {
    *r.count = count + 1;
  }
  return 1;
}

// sample_sequence (:100-118) given the drawn start position.  Returns the (possibly
// shifted) raw start; *len = number of frames, stopping after the first terminal.
UNREAL_HD int ring_sequence(const RingRef& r, int start, int seq_len, int* len) {
  if (frame_terminal(ring_at_raw(r, start))) start += 1;
  int n = 0;
  for (int i = 0; i < seq_len; ++i) {
    ++n;
    if (frame_terminal(ring_at_raw(r, start + i))) break;
  }
  *len = n;
  return start;
}

// rank-select: absolute index of the k-th (0-based) eligible frame whose reward sign matches
// (:137-142 `self._neg_reward_indices[index]` / `_pos_reward_indices[index]`).
UNREAL_HD int64_t ring_select(const RingRef& r, bool from_neg, int k) {
  const int64_t top = *r.top;
  int64_t lo = top + 3;
  if (lo < 3) lo = 3;
  const int64_t hi = top + *r.count - 1;
  for (int64_t a = lo; a <= hi; ++a) {
    const bool pos = frame_reward(ring_at_abs(r, a)) > 0;
    if (pos != from_neg) {
      if (k == 0) return a;
      --k;
    }
  }
  return -1;
}

}  // namespace unreal
