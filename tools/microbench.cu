// microbench.cu -- design-decision measurements on the B200 (developer tool, not product):
// shared-memory atomics vs ballot multisplit vs match.any, per-SM throughput at full occupancy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

template <int BITS>
__device__ __forceinline__ uint32_t match_digit(uint32_t digit)
{
	uint32_t peers = 0xffffffffu;
#pragma unroll
	for (int b = 0; b < BITS; ++b) {
		const bool bit = (digit >> b) & 1u;
		const uint32_t votes = __ballot_sync(0xffffffffu, bit);
		peers &= bit ? votes : ~votes;
	}
	return peers;
}

constexpr int ITERS = 4096;

// mode 0: smem atomicAdd (return value used), NB bins spread
// mode 1: smem atomicAdd no return
// mode 2: ballot match (BITS) + leader LDS/STS on per-warp counters
// mode 3: __match_any_sync + leader LDS/STS
// mode 4: ballot match only (no smem)
template <int MODE, int BITS>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
	constexpr int NB = 1 << BITS;
	extern __shared__ uint32_t sm[];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < NB * 8; i += 256) sm[i] = 0;
	__syncthreads();
	uint32_t *mine = sm + (MODE >= 2 ? warp * NB : 0);
	uint64_t x = mix64(seed + blockIdx.x * 256 + tid);
	uint32_t acc = 0;
	for (int it = 0; it < ITERS; it += 4) {
		x = mix64(x);
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const uint32_t d = uint32_t(x >> (u * 16)) & (NB - 1);
			if (MODE == 0) acc += atomicAdd(&sm[d], 1u);
			if (MODE == 1) atomicAdd(&sm[d], 1u);
			if (MODE == 2 || MODE == 3) {
				const uint32_t peers = MODE == 2 ? match_digit<BITS>(d) : __match_any_sync(0xffffffffu, d);
				const uint32_t leader = __ffs(peers) - 1;
				uint32_t before = 0;
				if (lane == leader) {
					before = mine[d];
					mine[d] = before + __popc(peers);
				}
				before = __shfl_sync(0xffffffffu, before, leader);
				acc += before + __popc(peers & ((1u << lane) - 1));
				__syncwarp();
			}
			if (MODE == 4) acc += __popc(match_digit<BITS>(d));
		}
	}
	if (acc == 0x12345678) out[0] = acc;
	if (tid == 0 && blockIdx.x == 0) out[1] = sm[5];
}

template <int MODE, int BITS>
void run(const char *name, int blocks_per_sm)
{
	uint32_t *out;
	cudaMalloc(&out, 64);
	const int sms = 148;
	const size_t smem = (size_t(1) << BITS) * 8 * 4;
	cudaFuncSetAttribute(k<MODE, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	k<MODE, BITS><<<sms * blocks_per_sm, 256, smem>>>(out, 1);
	cudaEventRecord(a);
	k<MODE, BITS><<<sms * blocks_per_sm, 256, smem>>>(out, 2);
	cudaEventRecord(b);
	cudaDeviceSynchronize();
	float ms;
	cudaEventElapsedTime(&ms, a, b);
	const double items = double(sms) * blocks_per_sm * 256 * ITERS;
	printf("%-34s bits=%2d blocks/SM=%d  %8.3f ms  %7.2f items/ns  %6.2f items/clk/SM @1.965GHz  (%s)\n",
	       name, BITS, blocks_per_sm, ms, items / ms / 1e6, items / ms / 1e6 / 148 / 1.965,
	       cudaGetErrorString(cudaGetLastError()));
	cudaFree(out);
}

int main()
{
	run<0, 8>("smem atomicAdd ret", 4);
	run<0, 8>("smem atomicAdd ret", 8);
	run<1, 8>("smem atomicAdd noret", 8);
	run<0, 12>("smem atomicAdd ret", 4);
	run<1, 12>("smem atomicAdd noret", 4);
	run<2, 8>("ballot match + LDS/STS", 4);
	run<2, 8>("ballot match + LDS/STS", 8);
	run<3, 8>("match.any + LDS/STS", 8);
	run<4, 8>("ballot match only", 8);
	run<2, 10>("ballot match + LDS/STS", 4);
	run<3, 10>("match.any + LDS/STS", 4);
	run<4, 10>("ballot match only", 4);
	run<4, 11>("ballot match only", 4);
	return 0;
}
