// K7: per-group confusion counts.
//
// Replaces the per-instance Python loops of the reference (tone_bias_test.py:207-234 builds one
// dict per image, :240-272 partitions them into TP/TN/FP/FN dicts, :283-289 filters by group,
// and only len() of each set is ever used) by one pass over three small byte arrays:
//
//   counts[a][g][label][pred] += 1     for every instance and every attribute a whose group id
//                                      g = groups[a][i] is a real group (g < n_groups)
//
// Instances whose group id is out of range are in no group of that attribute -- exactly what
// `filter(instances, feature, value)` does with NaN / unknown values.  Integer adds commute, so
// the result is bit-exact for any launch geometry and any sharding across GPUs.
//
// Memory-bound at <= (2 + n_attr) bytes per instance.  Each warp aggregates equal bins with
// __match_any_sync (one shared-memory atomic per distinct bin per warp), each CTA keeps a
// private int32 histogram in shared memory and flushes it once with 64-bit global atomics.
#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int COUNT_MAX_BINS = 1024;  // n_attr * n_groups * 4
constexpr int COUNT_THREADS = 256;

// Adds one instance per lane to the CTA histogram; all 32 lanes must call (inactive: bin < 0).
__device__ __forceinline__ void warp_aggregate_add(int* hist, int bin) {
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (bin >= 0 && (peers & ((1u << lane_id()) - 1u)) == 0) {  // lowest lane of each bin group
    atomicAdd(&hist[bin], __popc(peers));
  }
}

__global__ void __launch_bounds__(COUNT_THREADS)
confusion_counts_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ label,
                        const uint8_t* __restrict__ groups, long long n, long long groups_stride, int n_attr,
                        int n_groups, unsigned long long* __restrict__ counts) {
  __shared__ int hist[COUNT_MAX_BINS];
  const int bins = n_attr * n_groups * 4;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0;
  __syncthreads();

  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_round = (n + 31) / 32 * 32;  // keep warps converged for match_any
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool live = i < n;
    int cell = 0;
    if (live) cell = ((label[i] != 0) ? 2 : 0) | ((pred[i] != 0) ? 1 : 0);
    for (int a = 0; a < n_attr; ++a) {
      int bin = -1;
      if (live) {
        const int g = groups[(long long)a * groups_stride + i];
        if (g < n_groups) bin = (a * n_groups + g) * 4 + cell;
      }
      warp_aggregate_add(hist, bin);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    const int v = hist[i];
    if (v != 0) atomicAdd(&counts[i], (unsigned long long)v);
  }
}

}  // namespace sia

extern "C" int sia_confusion_counts(const uint8_t* pred, const uint8_t* label, const uint8_t* groups, long long n,
                                    long long groups_stride, int n_attr, int n_groups, long long* counts,
                                    void* stream) {
  using namespace sia;
  SIA_REQUIRE(counts != nullptr && n >= 0 && n_attr >= 1 && n_groups >= 1);
  SIA_REQUIRE(n_attr * n_groups * 4 <= COUNT_MAX_BINS && groups_stride >= n);
  if (n == 0) return 0;  // empty shard: nothing to add
  SIA_REQUIRE(pred && label && groups);
  // int32 per-CTA histogram: bound the instances one CTA can see well below 2^31
  long long blocks = (n + COUNT_THREADS * 8 - 1) / (COUNT_THREADS * 8);
  const long long max_blocks = (long long)sm_count() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  const long long per_block = (n + blocks - 1) / blocks;
  if (per_block > (1ll << 30)) return SIA_E_UNSUPPORTED;
  confusion_counts_kernel<<<(unsigned)blocks, COUNT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, label, groups, n, groups_stride, n_attr, n_groups, reinterpret_cast<unsigned long long*>(counts));
  return launch_status();
}
