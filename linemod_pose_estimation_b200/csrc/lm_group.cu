// lm_group.cu -- several GPUs behind ONE C++ caller: the multi-GPU handle of SURVEY 8b ("owns R devices").
//
// The reference's caller is one process that makes one call (rgbdDetector::linemod_detection,
// /root/reference/src/rgbdDetector.cpp:31-34, from the service loop of linemod_ensenso_detect_3_mult_detect_service.cpp).
// A group clones a detector onto N devices and drives every device from its own worker thread:
//
//   LM_GROUP_FRAMES     every device holds all templates; the frames of a batch are dealt out in launch sets (chunks of
//                       "batch_frames" frames, round robin).  Each device pulls its frames straight from the caller's host
//                       memory over its own PCIe link and returns finished match lists: no exchange between devices at
//                       all (SURVEY 8e option C) -- the mode that scales the end-to-end frame rate.
//   LM_GROUP_TEMPLATES  the north-star layout: templates sharded by canonical index (lm_set_shard), every device sees
//                       every frame, and the survivors of all shards are merged before the reference's sort + unique
//                       (std::unique acts across templates, App. D-7).  For the single-frame latency case and for template
//                       sets that outgrow one device.
//
//   grid (lm_group_create_grid)  both at once: the devices form N / S sets of S template shards; frames are dealt out to the
//                       sets, the S devices of a set see the set's frames and their survivors are merged.  S = 1 is the
//                       frames mode, S = N the templates mode; in between, a template set too large (or a latency target
//                       too tight) for one GPU still scales its frame rate with the number of sets.
//
// In one process the host is both the source of the frames and the consumer of the matches, so neither mode has a
// device-to-device exchange step: every device copies in over its own link and its (few hundred bytes of) survivors go
// straight back to the host thread that merges them.  The NCCL exchange (frame broadcast + survivor all-gather over
// NVLink) belongs to the one-process-per-GPU deployment and lives in linemod_pose_estimation_b200/sharding.py.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "lm_detector_internal.hpp"

struct lm_group {
  int mode = LM_GROUP_FRAMES;
  int shards = 1;  // template shards per set of devices; the group has det.size() / shards sets
  std::vector<lm_detector*> det;
  std::vector<int> device;
  std::vector<std::thread> worker;
  std::mutex mu;
  std::condition_variable wake, done;
  const std::function<int(int)>* job = nullptr;
  unsigned long epoch = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rc;
  std::vector<std::string> err;

  void loop(int i) {
    cudaSetDevice(device[i]);  // the member handle binds to the thread's current device on first use
    unsigned long seen = 0;
    for (;;) {
      const std::function<int(int)>* fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu);
        wake.wait(lk, [&]() { return epoch != seen; });
        seen = epoch;
        if (stop) return;
        fn = job;
      }
      int r = LM_OK;
      std::string e;
      try {
        r = (*fn)(i);
        if (r != LM_OK) e = lm_last_error();   // thread-local: copy it before it is lost with this thread's next call
      } catch (const std::exception& ex) { r = LM_E_INVALID; e = ex.what(); }
      {
        std::lock_guard<std::mutex> lk(mu);
        rc[i] = r; err[i] = e;
        if (--pending == 0) done.notify_one();
      }
    }
  }
  // fn(i) on the worker of every device; first failure wins
  int run(const std::function<int(int)>& fn) {
    {
      std::lock_guard<std::mutex> lk(mu);
      job = &fn; pending = (int)det.size(); ++epoch;
      for (size_t i = 0; i < rc.size(); ++i) { rc[i] = LM_OK; err[i].clear(); }
    }
    wake.notify_all();
    std::unique_lock<std::mutex> lk(mu);
    done.wait(lk, [this]() { return pending == 0; });
    job = nullptr;
    for (size_t i = 0; i < rc.size(); ++i)
      if (rc[i] != LM_OK) return lm_fail(rc[i], "device %d: %s", device[i], err[i].c_str());
    return LM_OK;
  }
};

extern "C" {

int lm_group_create_grid(const lm_detector* prototype, const int* devices, int n_devices, int template_shards, lm_group** out) {
  if (!prototype || !devices || !out || n_devices < 1) return lm_fail(LM_E_INVALID, "bad argument");
  if (template_shards < 1 || n_devices % template_shards != 0)
    return lm_fail(LM_E_INVALID, "template_shards (%d) must divide the number of devices (%d)", template_shards, n_devices);
  const int mode = template_shards == 1 ? LM_GROUP_FRAMES : LM_GROUP_TEMPLATES;
  *out = nullptr;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
    cudaGetLastError();
    return lm_fail(LM_E_CUDA, "no CUDA device available; this library has no CPU path");
  }
  for (int i = 0; i < n_devices; ++i)
    if (devices[i] < 0 || devices[i] >= have) return lm_fail(LM_E_INVALID, "device %d does not exist (%d devices)", devices[i], have);
  lm_group* g = new lm_group();
  g->mode = mode;
  g->shards = template_shards;
  g->device.assign(devices, devices + n_devices);
  g->rc.assign((size_t)n_devices, LM_OK);
  g->err.assign((size_t)n_devices, std::string());
  for (int i = 0; i < n_devices; ++i) {
    lm_detector* d = lm_internal_clone(prototype);
    if (template_shards > 1) lm_set_shard(d, i % template_shards, template_shards);   // device i = shard i % S of set i / S
    g->det.push_back(d);
  }
  for (int i = 0; i < n_devices; ++i) g->worker.emplace_back([g, i]() { g->loop(i); });
  *out = g;
  return LM_OK;
}

int lm_group_create(const lm_detector* prototype, const int* devices, int n_devices, int mode, lm_group** out) {
  if (mode != LM_GROUP_FRAMES && mode != LM_GROUP_TEMPLATES) return lm_fail(LM_E_INVALID, "unknown group mode %d", mode);
  return lm_group_create_grid(prototype, devices, n_devices, mode == LM_GROUP_FRAMES ? 1 : n_devices, out);
}

void lm_group_destroy(lm_group* g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->stop = true; ++g->epoch;
  }
  g->wake.notify_all();
  for (auto& t : g->worker) t.join();
  for (lm_detector* d : g->det) lm_destroy(d);
  delete g;
}

int lm_group_size(const lm_group* g) { return g ? (int)g->det.size() : 0; }
int lm_group_mode(const lm_group* g) { return g ? g->mode : -1; }
lm_detector* lm_group_member(lm_group* g, int i) {
  if (!g || i < 0 || i >= (int)g->det.size()) { lm_fail(LM_E_INVALID, "member out of range"); return nullptr; }
  return g->det[(size_t)i];
}
int lm_group_set_option(lm_group* g, const char* key, int value) {
  if (!g) return lm_fail(LM_E_INVALID, "NULL argument");
  for (lm_detector* d : g->det) {
    int rc = lm_set_option(d, key, value);
    if (rc != LM_OK) return rc;
  }
  return LM_OK;
}

int lm_group_match_batch_multi(lm_group* g, const lm_image* sources, int n_frames, int n_sources, const lm_query* queries,
                               int n_queries, lm_match_rec** out_matches, size_t* out_offsets) {
  if (!g || !sources || !queries || !out_matches || !out_offsets || n_frames < 0 || n_queries < 1 || n_queries > LM_MAX_QUERIES)
    return lm_fail(LM_E_INVALID, "bad argument");
  *out_matches = nullptr;
  out_offsets[0] = 0;
  const int N = (int)g->det.size(), M = n_sources, Q = n_queries, S = g->shards, SETS = N / S;
  std::vector<lm_match_rec> all;
  // launch sets dealt round robin over the sets of devices: set k takes frames [c * F, (c + 1) * F) for c % SETS == k
  const int F = std::max(1, std::min(g->det[0]->batch_frames, LM_MAX_BATCH));
  std::vector<std::vector<lm_image> > mine((size_t)SETS);
  std::vector<std::vector<int> > frame_of((size_t)SETS);
  for (int f = 0; f < n_frames; ++f) {
    const int k = (f / F) % SETS;
    for (int m = 0; m < M; ++m) mine[(size_t)k].push_back(sources[(size_t)f * M + m]);
    frame_of[(size_t)k].push_back(f);
  }
  std::vector<std::pair<int, int> > where((size_t)n_frames);  // frame -> (set, index within the set's frames)
  for (int k = 0; k < SETS; ++k)
    for (size_t j = 0; j < frame_of[(size_t)k].size(); ++j) where[(size_t)frame_of[(size_t)k][j]] = std::make_pair(k, (int)j);
  if (S == 1) {
    // every device holds all templates: finished lists per device, put back into frame order
    std::vector<lm_match_rec*> part((size_t)N, nullptr);
    std::vector<std::vector<size_t> > offs((size_t)N);
    int rc = g->run([&](int i) -> int {
      const int n = (int)frame_of[(size_t)i].size();
      offs[(size_t)i].assign((size_t)n * Q + 1, 0);
      if (n == 0) return LM_OK;
      return lm_match_batch_multi(g->det[(size_t)i], mine[(size_t)i].data(), n, M, queries, Q, &part[(size_t)i], offs[(size_t)i].data());
    });
    if (rc == LM_OK) {
      for (int f = 0; f < n_frames; ++f) {
        const int i = where[(size_t)f].first, k = where[(size_t)f].second;
        for (int q = 0; q < Q; ++q) {
          const size_t a = offs[(size_t)i][(size_t)k * Q + q], b = offs[(size_t)i][(size_t)k * Q + q + 1];
          all.insert(all.end(), part[(size_t)i] + a, part[(size_t)i] + b);
          out_offsets[(size_t)f * Q + q + 1] = all.size();
        }
      }
    }
    for (lm_match_rec* p : part) lm_free_matches(p);
    if (rc != LM_OK) return rc;
  } else {
    // every device of a set matches the set's frames against its template shard; the shards' survivors are merged per frame
    std::vector<std::vector<std::vector<lm_raw_match> > > raw((size_t)N);
    int rc = g->run([&](int i) -> int {
      const int k = i / S, n = (int)frame_of[(size_t)k].size();
      if (n == 0) { raw[(size_t)i].clear(); return LM_OK; }
      return lm_internal_match_batch_raw(g->det[(size_t)i], mine[(size_t)k].data(), n, M, queries, Q, &raw[(size_t)i]);
    });
    if (rc != LM_OK) return rc;
    const int levels = lm_pyramid_levels(g->det[0]);
    std::vector<lm_raw_match> merged;
    std::vector<lm_match_rec> out[LM_MAX_QUERIES];
    for (int f = 0; f < n_frames; ++f) {
      const int k = where[(size_t)f].first, j = where[(size_t)f].second;
      merged.clear();
      for (int i = k * S; i < (k + 1) * S; ++i) merged.insert(merged.end(), raw[(size_t)i][(size_t)j].begin(), raw[(size_t)i][(size_t)j].end());
      lm_internal_finalize(levels, merged, Q, out);
      for (int q = 0; q < Q; ++q) {
        all.insert(all.end(), out[q].begin(), out[q].end());
        out_offsets[(size_t)f * Q + q + 1] = all.size();
      }
    }
  }
  lm_match_rec* p = (lm_match_rec*)std::malloc(std::max<size_t>(1, all.size()) * sizeof(lm_match_rec));
  if (!p) return lm_fail(LM_E_INVALID, "out of host memory");
  if (!all.empty()) std::memcpy(p, all.data(), all.size() * sizeof(lm_match_rec));
  *out_matches = p;
  return LM_OK;
}

int lm_group_match(lm_group* g, const lm_image* sources, int n_sources, float threshold, const char* const* class_ids,
                   int n_ids, lm_match_rec** out_matches, size_t* out_n) {
  if (!out_n) return lm_fail(LM_E_INVALID, "NULL argument");
  *out_n = 0;
  lm_query q = {threshold, class_ids, n_ids};
  size_t offs[2] = {0, 0};
  int rc = lm_group_match_batch_multi(g, sources, 1, n_sources, &q, 1, out_matches, offs);
  if (rc == LM_OK) *out_n = offs[1];
  return rc;
}

}  // extern "C"
