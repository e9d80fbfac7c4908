// probe: 2D TMA tile load of rows taken two at a time (row stride 625 samples)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap pm, const void *gm, int c0, int c1, float *out)
{
	extern __shared__ __align__(1024) unsigned char sm[];
	float *buf = reinterpret_cast<float *>(sm);
	unsigned bar = smem_u32(sm + 4096);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2048u) : "memory");
		const void *tm = MODE == 0 ? (const void *)&pm : gm;
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(buf)),
			     "l"(tm), "r"(c0), "r"(c1), "r"(bar)
			     : "memory");
	}
	unsigned done = 0;
	int spins = 0;
	while (!done && spins < (1 << 22)) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(0u) : "memory");
		spins++;
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = done ? buf[i] : -1.0f;
}
int main()
{
	const int n = 64, stride = 625;
	std::vector<float> h((size_t)n * stride * 2);
	for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
	float *d, *o;
	cudaMalloc(&d, h.size() * 4);
	cudaMalloc(&o, 512 * 4);
	cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
	void *fn = nullptr;
	cudaDriverEntryPointQueryResult qr;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
	typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
				     const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	EncodeFn enc = (EncodeFn)fn;
	alignas(64) CUtensorMap tm;
	const cuuint64_t gdim[2] = { (cuuint64_t)4 * stride, (cuuint64_t)(n / 2) };
	const cuuint64_t gstr[1] = { (cuuint64_t)16 * stride };
	const cuuint32_t box[2] = { 32u, 16u }, est[2] = { 1u, 1u };
	for (int swz = 0; swz < 2; swz++) {
		CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
				 swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		printf("encode swz=%d -> %d (qr %d)\n", swz, (int)r, (int)qr);
		void *gm;
		cudaMalloc(&gm, 128);
		cudaMemcpy(gm, &tm, 128, cudaMemcpyHostToDevice);
		for (int mode = 0; mode < 2; mode++) {
			const int c0 = 2 * 208, c1 = 3;
			if (mode == 0) k<0><<<1, 64, 8192>>>(tm, gm, c0, c1, o);
			else k<1><<<1, 64, 8192>>>(tm, gm, c0, c1, o);
			cudaError_t e = cudaDeviceSynchronize();
			float ho[512];
			cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
			printf("swz %d mode %d: %s; row0: %.0f %.0f %.0f %.0f | row1: %.0f %.0f ; expect %.0f and %.0f\n", swz, mode, cudaGetErrorString(e), ho[0], ho[1], ho[2], ho[3],
			       ho[32], ho[33], (double)((size_t)(2 * c1) * stride * 2 + c0), (double)((size_t)(2 * (c1 + 1)) * stride * 2 + c0));
			if (e != cudaSuccess) return 1;
		}
	}
	return 0;
}
