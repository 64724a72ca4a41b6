// p2p_bench.cu -- what NVLink gives a kernel that stores into a peer GPU's memory.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/p2p_bench tools/p2p_bench.cu
//   ./tools/p2p_bench [GiB per direction]            (needs 2 GPUs with peer access)
//
// Measures, one direction (GPU0 -> GPU1) and both directions at once (the exchange of the
// sharded sort is bidirectional):
//   memcpy   cudaMemcpyPeerAsync (copy engines): the practical ceiling of the link
//   st8      kernel, every lane stores 8 bytes, warps store 256 contiguous bytes
//   st16     kernel, every lane stores 16 bytes
//   bulk     kernel, shared memory -> peer global by cp.async.bulk (TMA store), 4 KiB pieces
// Sources are local HBM reads; grid = resident blocks x SMs.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) st8_kernel(const uint64_t *src, uint64_t *dst, size_t n)
{
	for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
		dst[i] = src[i];
}

__global__ void __launch_bounds__(256) st16_kernel(const ulonglong2 *src, ulonglong2 *dst, size_t n2)
{
	for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n2; i += size_t(gridDim.x) * blockDim.x)
		dst[i] = src[i];
}

constexpr int PIECE = 4096;          // bytes per bulk store
constexpr int PIECES = 8;            // pieces per block iteration (32 KiB of shared memory)

__device__ __forceinline__ uint32_t saddr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(256) bulk_kernel(const ulonglong2 *src, char *dst, size_t bytes)
{
	extern __shared__ __align__(128) unsigned char sm[];
	const size_t chunk = size_t(PIECE) * PIECES;
	for (size_t off = blockIdx.x * chunk; off + chunk <= bytes; off += size_t(gridDim.x) * chunk) {
		// wait until the previous bulk stores have finished READING shared memory
		if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
		__syncthreads();
		const ulonglong2 *s = reinterpret_cast<const ulonglong2 *>(reinterpret_cast<const char *>(src) + off);
		ulonglong2 *d = reinterpret_cast<ulonglong2 *>(sm);
		for (int i = threadIdx.x; i < int(chunk / 16); i += blockDim.x) d[i] = s[i];
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncthreads();
		if (threadIdx.x < PIECES) {
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
				     :: "l"(dst + off + size_t(threadIdx.x) * PIECE), "r"(saddr(sm + threadIdx.x * PIECE)), "r"(PIECE)
				     : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
	}
	if (threadIdx.x < PIECES) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Side {
	int dev;
	cudaStream_t st;
	cudaEvent_t a, b;
	char *local, *remote;     // remote = buffer on the other GPU
};

template <class F>
static void run(const char *name, Side *s, int sides, size_t bytes, F launch)
{
	for (int rep = 0; rep < 3; ++rep) {
		for (int i = 0; i < sides; ++i) { CK(cudaSetDevice(s[i].dev)); CK(cudaDeviceSynchronize()); }
		for (int i = 0; i < sides; ++i) { CK(cudaSetDevice(s[i].dev)); CK(cudaEventRecord(s[i].a, s[i].st)); launch(s[i]); CK(cudaEventRecord(s[i].b, s[i].st)); }
		for (int i = 0; i < sides; ++i) { CK(cudaSetDevice(s[i].dev)); CK(cudaStreamSynchronize(s[i].st)); CK(cudaGetLastError()); }
		if (rep == 2) {
			float worst = 0;
			for (int i = 0; i < sides; ++i) { float ms; CK(cudaEventElapsedTime(&ms, s[i].a, s[i].b)); worst = ms > worst ? ms : worst; }
			printf("%-8s %s: %8.3f ms  %7.1f GB/s per direction\n", name, sides == 1 ? "one direction  " : "both directions",
			       worst, bytes / (worst * 1e-3) / 1e9);
		}
	}
}

int main(int argc, char **argv)
{
	const size_t bytes = size_t(argc > 1 ? atof(argv[1]) * 1024 : 4096) << 20;
	int n = 0;
	CK(cudaGetDeviceCount(&n));
	if (n < 2) { printf("need 2 GPUs\n"); return 0; }
	int can = 0;
	CK(cudaDeviceCanAccessPeer(&can, 0, 1));
	if (!can) { printf("no peer access between GPU 0 and 1\n"); return 0; }
	Side s[2];
	char *buf[2][2];
	int sms = 0;
	for (int d = 0; d < 2; ++d) {
		CK(cudaSetDevice(d));
		CK(cudaDeviceEnablePeerAccess(1 - d, 0));
		CK(cudaMalloc(&buf[d][0], bytes));
		CK(cudaMalloc(&buf[d][1], bytes));
		CK(cudaMemset(buf[d][0], d + 1, bytes));
		CK(cudaStreamCreate(&s[d].st));
		CK(cudaEventCreate(&s[d].a));
		CK(cudaEventCreate(&s[d].b));
		CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d));
		CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIECE * PIECES));
	}
	for (int d = 0; d < 2; ++d) { s[d].dev = d; s[d].local = buf[d][0]; s[d].remote = buf[1 - d][1]; }
	printf("%.1f GiB per direction, %d SMs\n", bytes / double(1 << 30), sms);
	for (int sides = 1; sides <= 2; ++sides) {
		run("memcpy", s, sides, bytes, [&](Side &x) { CK(cudaMemcpyPeerAsync(x.remote, 1 - x.dev, x.local, x.dev, bytes, x.st)); });
		for (int bps : {2, 4, 8}) {
			char nm[32];
			snprintf(nm, sizeof nm, "st8 x%d", bps);
			run(nm, s, sides, bytes, [&](Side &x) { st8_kernel<<<sms * bps, 256, 0, x.st>>>((const uint64_t *) x.local, (uint64_t *) x.remote, bytes / 8); });
			snprintf(nm, sizeof nm, "st16 x%d", bps);
			run(nm, s, sides, bytes, [&](Side &x) { st16_kernel<<<sms * bps, 256, 0, x.st>>>((const ulonglong2 *) x.local, (ulonglong2 *) x.remote, bytes / 16); });
		}
		for (int bps : {1, 2, 4}) {
			char nm[32];
			snprintf(nm, sizeof nm, "bulk x%d", bps);
			run(nm, s, sides, bytes, [&](Side &x) { bulk_kernel<<<sms * bps, 256, PIECE * PIECES, x.st>>>((const ulonglong2 *) x.local, x.remote, bytes); });
		}
	}
	// sanity: the last transfer really landed
	CK(cudaSetDevice(1));
	unsigned char probe[2];
	CK(cudaMemcpy(probe, buf[1][1] + bytes / 2, 1, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(probe + 1, buf[1][1] + bytes - PIECE * PIECES - 1, 1, cudaMemcpyDeviceToHost));
	printf("probe %d %d (expect 1 1)\n", probe[0], probe[1]);
	return 0;
}
