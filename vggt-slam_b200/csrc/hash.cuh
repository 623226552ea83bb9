// hash.cuh -- the global voxel hash (device side): packed key -> dense voxel id.
#pragma once
#include "state.cuh"

namespace vsm {

struct GlobalStore {
  unsigned long long* gkeys;
  int32_t* gids;
  uint64_t gmask;
  unsigned long long* vkey;
  uint32_t* vcount;
  uint32_t* n_vox;
  uint32_t vcap;
};

// returns the dense voxel id of `key`, inserting it if absent; -1 on overflow (flagged)
__device__ __forceinline__ int global_find_or_insert(const GlobalStore& g, unsigned long long key, uint32_t* err) {
  uint64_t h = mix64(key) & g.gmask;
  for (uint64_t probes = 0; probes <= g.gmask; ++probes) {
    unsigned long long cur = g.gkeys[h];
    if (cur == kEmptyKey) {
      cur = atomicCAS(&g.gkeys[h], kEmptyKey, key);
      if (cur == kEmptyKey) {
        const uint32_t id = atomicAdd(g.n_vox, 1u);
        if (id >= g.vcap) {
          atomicAdd(err, 1u);
          reinterpret_cast<volatile int32_t*>(g.gids)[h] = -2;  // release waiters; callers treat < 0 as failure
          return -1;
        }
        g.vkey[id] = key;
        __threadfence();
        reinterpret_cast<volatile int32_t*>(g.gids)[h] = (int32_t)id;
        return (int)id;
      }
    }
    if (cur == key) {
      int id;
      // the claimer publishes the id right after its CAS; within one launch keys are distinct, so this
      // spins only if two launches raced on different streams
      while ((id = reinterpret_cast<volatile int32_t*>(g.gids)[h]) == -1) {
      }
      return id;
    }
    h = (h + 1) & g.gmask;
  }
  atomicAdd(err, 1u);
  return -1;
}


static inline GlobalStore global_store(vsm_map* m) {
  GlobalStore g;
  g.gkeys = m->gkeys.as<unsigned long long>();
  g.gids = m->gids.as<int32_t>();
  g.gmask = m->gcap - 1;
  g.vkey = m->vkey.as<unsigned long long>();
  g.vcount = m->vcount.as<uint32_t>();
  g.n_vox = m->d_n_vox.as<uint32_t>();
  g.vcap = (uint32_t)m->vcap;
  return g;
}


}  // namespace vsm
