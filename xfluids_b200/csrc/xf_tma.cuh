// xf_tma.cuh -- TMA (cp.async.bulk.tensor) and mbarrier primitives for sm_100a as inline PTX (PTX ISA 8.x; SASS: UTMALDG / SYNCS).
// The tensor maps are encoded on the host (xf_capi.cu, cuTensorMapEncodeTiled) and passed as __grid_constant__ kernel parameters.
#pragma once
#include <cuda.h>

__device__ __forceinline__ unsigned xf_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xf_mbar_init(unsigned long long *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xf_smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void xf_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xf_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xf_mbar_wait(unsigned long long *bar, unsigned parity)
{
	unsigned ok;
	do
	{
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
					 : "=r"(ok)
					 : "r"(xf_smem_u32(bar)), "r"(parity)
					 : "memory");
	} while (!ok);
}
__device__ __forceinline__ void xf_tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(xf_smem_u32(dst)),
				 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(xf_smem_u32(bar))
				 : "memory");
}

