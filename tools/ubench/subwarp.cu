// micro-probe: do redux.sync / ballot / shfl work with per-8-lane member masks when the four groups of a warp execute them converged?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k( uint32_t* out )
{
	const uint32_t lane = threadIdx.x & 31, g = lane >> 3;
	const uint32_t gmask = 0xffu << (g * 8);
	const uint32_t v = 1u << lane;
	const uint32_t r = __reduce_or_sync( gmask, v );
	const uint32_t b = __ballot_sync( gmask, (lane & 1) != 0 );
	const uint32_t mn = __reduce_min_sync( gmask, lane * 7u % 13u );
	const uint32_t s = __shfl_sync( gmask, lane, (g * 8) + 3 );
	out[threadIdx.x * 4 + 0] = r, out[threadIdx.x * 4 + 1] = b, out[threadIdx.x * 4 + 2] = mn, out[threadIdx.x * 4 + 3] = s;
}
int main()
{
	uint32_t* d; uint32_t h[128];
	cudaMalloc( &d, sizeof( h ) );
	k<<<1, 32>>>( d );
	cudaError_t e = cudaDeviceSynchronize();
	printf( "sync: %s\n", cudaGetErrorString( e ) );
	cudaMemcpy( h, d, sizeof( h ), cudaMemcpyDeviceToHost );
	int bad = 0;
	for (int l = 0; l < 32; l++)
	{
		const uint32_t g = l >> 3, gm = 0xffu << (g * 8);
		uint32_t mn = 99; for (int j = 0; j < 8; j++) { uint32_t x = (g * 8 + j) * 7u % 13u; if (x < mn) mn = x; }
		if (h[l * 4] != gm || h[l * 4 + 1] != (0xaaaaaaaau & gm) || h[l * 4 + 2] != mn || h[l * 4 + 3] != g * 8 + 3) bad++, printf( "lane %d: or %08x ballot %08x min %u shfl %u\n", l, h[l * 4], h[l * 4 + 1], h[l * 4 + 2], h[l * 4 + 3] );
	}
	printf( "subwarp redux/ballot/shfl with per-group masks: %s\n", bad ? "MISMATCH" : "ok" );
	return bad != 0;
}
