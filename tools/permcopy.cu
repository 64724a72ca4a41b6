// permcopy.cu -- memory-system ceiling of the scatter write pattern (developer tool).
// A tile of 4096 pairs is read sequentially and written as NB runs of L = 4096/NB pairs,
// run b going to region b of the output (exactly what one radix pass does for uniform
// keys), with no ranking and no shared memory.  Reports read+write GB/s per L.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/permcopy tools/permcopy.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int TILE = 4096;

template <int L, bool PAIRS, int VEC>
__global__ void __launch_bounds__(256) permcopy(const uint64_t *__restrict__ k_in, const uint64_t *__restrict__ r_in,
						uint64_t *__restrict__ k_out, uint64_t *__restrict__ r_out,
						uint32_t ntiles, uint32_t region)
{
	for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const uint64_t base = uint64_t(t) * TILE;
		if (VEC == 1) {
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const uint32_t i = j * 256 + threadIdx.x;
				const uint32_t b = i / L, o = i % L;
				const uint64_t dst = uint64_t(b) * region + uint64_t(t) * L + o;
				k_out[dst] = k_in[base + i];
				if (PAIRS) r_out[dst] = r_in[base + i];
			}
		} else {
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const uint32_t i = (j * 256 + threadIdx.x) * 2;
				const uint32_t b = i / L, o = i % L;
				const uint64_t dst = uint64_t(b) * region + uint64_t(t) * L + o;
				*reinterpret_cast<ulonglong2 *>(k_out + dst) = *reinterpret_cast<const ulonglong2 *>(k_in + base + i);
				if (PAIRS) *reinterpret_cast<ulonglong2 *>(r_out + dst) = *reinterpret_cast<const ulonglong2 *>(r_in + base + i);
			}
		}
	}
}

template <int L, bool PAIRS, int VEC>
void run(uint64_t *a, uint64_t *b, uint64_t *c, uint64_t *d, uint64_t n, int blocks)
{
	const uint32_t ntiles = n / TILE;
	const uint32_t nb = TILE / L;
	const uint32_t region = n / nb;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	permcopy<L, PAIRS, VEC><<<blocks, 256>>>(a, b, c, d, ntiles, region);
	cudaEventRecord(e0);
	permcopy<L, PAIRS, VEC><<<blocks, 256>>>(a, b, c, d, ntiles, region);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double bytes = double(n) * 8 * 2 * (PAIRS ? 2 : 1);
	printf("run %4d B (%4d bins) %s vec%d blocks=%5d: %7.3f ms  %7.1f GB/s  (%s)\n", L * 8, nb,
	       PAIRS ? "pairs" : "keys ", VEC, blocks, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv)
{
	const uint64_t n = 1ull << (argc > 1 ? atoi(argv[1]) : 29);
	uint64_t *a, *b, *c, *d;
	cudaMalloc(&a, n * 8);
	cudaMalloc(&b, n * 8);
	cudaMalloc(&c, n * 8);
	cudaMalloc(&d, n * 8);
	cudaMemset(a, 1, n * 8);
	cudaMemset(b, 2, n * 8);
	for (int blocks : {148 * 2, 148 * 4, 148 * 8}) {
		run<4096, true, 1>(a, b, c, d, n, blocks);   // plain copy
		run<4096, true, 2>(a, b, c, d, n, blocks);
		run<128, true, 1>(a, b, c, d, n, blocks);
		run<64, true, 1>(a, b, c, d, n, blocks);
		run<32, true, 1>(a, b, c, d, n, blocks);
		run<16, true, 1>(a, b, c, d, n, blocks);
		run<16, true, 2>(a, b, c, d, n, blocks);
		run<16, false, 1>(a, b, c, d, n, blocks);
		run<8, true, 1>(a, b, c, d, n, blocks);
		run<4, true, 1>(a, b, c, d, n, blocks);
		run<2, true, 1>(a, b, c, d, n, blocks);
	}
	return 0;
}
