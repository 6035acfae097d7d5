// bulk_copy.cuh -- sm_100a asynchronous bulk copies (TMA engine, SASS: UBLKCP) with mbarrier completion, and the
// predicated FMA used by the fused Gauss-Newton kernel.
//
// The keyframe maps are flat fp32 arrays (pitch == cols * 4), so a warp's slice of a tile is one contiguous
// 512-byte segment per map: a 1-D bulk copy needs no tensor map and is issued by a single lane.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rgbid {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// global -> shared bulk copy; bytes, both addresses multiples of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// The same with an L2 eviction policy.  The fused kernel streams each keyframe map exactly once per launch (236 MB per
// launch against a 126 MB L2): marked evict-first they pass through without displacing the current-frame maps its
// texture gathers re-read every iteration, the solver state, the partial sums and the kernels' own instructions.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// programmatic dependent launch (no-ops when the kernel was launched without the attribute): wait = every prerequisite
// grid has completed and its writes are visible; launch_dependents = this CTA no longer holds back the launch of the
// next kernel's CTAs (they start once every CTA of this grid has said so or exited)
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// true in exactly one lane of a converged warp (the compiler then knows a single thread issues the bulk copies)
__device__ __forceinline__ bool elect_one()
{
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// d += a * b where flag != 0 (predicated, not branched: invalid pixels carry NaN operands that must not be
// accumulated, and a branch per pixel and constraint would serialise the four pixels of a thread)
__device__ __forceinline__ void pfma(float& d, float a, float b, int flag)
{
  asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\n@p fma.rn.f32 %0, %1, %2, %0;\n}" : "+f"(d) : "f"(a), "f"(b), "r"(flag));
}

}  // namespace rgbid
