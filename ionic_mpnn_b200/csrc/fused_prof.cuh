// Phase timing of the planned fused forward kernels (profiling builds only: -DF6_PHASE_PROF; tools/fused_prof2.py).
#pragma once
namespace imp {
// Optional phase timing (tools/fused_prof2.py builds with -DF6_PHASE_PROF): clock() deltas of thread 0 (class 0) and thread 96
// (class 1) of every context, summed per phase in shared memory and flushed to a device symbol at the end of the kernel.
#ifdef F6_PHASE_PROF
static __device__ unsigned long long f6_prof_total[2][16];
#define F6_PROF_DECL                                               \
  __shared__ unsigned int sprof[2][16];                            \
  if (tid < 32) sprof[tid >> 4][tid & 15] = 0u;                    \
  const int prof_cls = (t == 0) ? 0 : (t == 96) ? 1 : -1;          \
  unsigned int prof_last = (unsigned int)clock()
#define F6_PROF(i)                                                  \
  do {                                                             \
    if (prof_cls >= 0) {                                           \
      const unsigned int now_ = (unsigned int)clock();             \
      atomicAdd(&sprof[prof_cls][i], now_ - prof_last);            \
      prof_last = now_;                                            \
    }                                                              \
  } while (0)
#define F6_PROF_FLUSH \
  if (tid < 32) atomicAdd(&f6_prof_total[tid >> 4][tid & 15], (unsigned long long)sprof[tid >> 4][tid & 15])
#else
#define F6_PROF_DECL
#define F6_PROF(i)
#define F6_PROF_FLUSH
#endif

}  // namespace imp
