// micro-probe: issue rates of the instructions the CWBVH node step is made of (sm_100a): FFMA vs FFMA2, FADD vs FADD2, PRMT, FMNMX3, I2F.U8,
// and FFMA+PRMT interleaved.  Reports warp-instructions per clock per SM at full occupancy (148 x 8 CTAs x 256 threads).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define U 8
template <int MODE> __global__ void __launch_bounds__( 256 ) k( float* out, const float s, const uint32_t w )
{
	float a[U]; float2 p[U]; uint32_t q[U];
	for (int i = 0; i < U; i++) a[i] = threadIdx.x * 0.001f + i, p[i] = make_float2( a[i], a[i] + 1 ), q[i] = threadIdx.x * 2654435761u + i;
	const float2 s2 = make_float2( s, s * 1.0001f ), t2 = make_float2( 0.5f, 0.25f );
	for (int it = 0; it < ITERS; it++)
	{
		#pragma unroll
		for (int i = 0; i < U; i++)
		{
			if (MODE == 0) a[i] = __fmaf_rn( a[i], s, 0.5f + a[(i + 1) % U] * 0 );          // FFMA (3-reg form)
			if (MODE == 1) p[i] = __ffma2_rn( p[i], s2, t2 );                              // FFMA2
			if (MODE == 2) a[i] = __fadd_rn( a[i], s );                                    // FADD
			if (MODE == 3) p[i] = __fadd2_rn( p[i], s2 );                                  // FADD2
			if (MODE == 4) q[i] = __byte_perm( q[i], w, 0x7650u | (q[i] & 3u) );           // PRMT (+LOP)
			if (MODE == 5) a[i] = fmaxf( fmaxf( a[i], s ), a[(i + 3) % U] );               // FMNMX3
			if (MODE == 6) a[i] = (float)(uint8_t)(__float_as_uint( a[i] ) >> 8);           // I2F.U8
			if (MODE == 7) { a[i] = __fmaf_rn( a[i], s, 0.5f ); q[i] = __byte_perm( q[i], w, 0x7651u ); } // FFMA + PRMT pair
			if (MODE == 8) { p[i] = __ffma2_rn( p[i], s2, t2 ); q[i] = __byte_perm( q[i], w, 0x7651u ); q[(i + 1) % U] ^= q[i] >> 3; } // FFMA2 + 2 alu
		}
	}
	float r = 0; for (int i = 0; i < U; i++) r += a[i] + p[i].x + p[i].y + __uint_as_float( q[i] & 0x3fffffffu );
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run( const char* name, int per_iter, float* d )
{
	cudaEvent_t e0, e1; cudaEventCreate( &e0 ), cudaEventCreate( &e1 );
	const int grid = 148 * 8;
	k<MODE><<<grid, 256>>>( d, 1.0001f, 0x4b000000u );
	cudaEventRecord( e0 );
	k<MODE><<<grid, 256>>>( d, 1.0001f, 0x4b000000u );
	cudaEventRecord( e1 ); cudaEventSynchronize( e1 );
	float ms; cudaEventElapsedTime( &ms, e0, e1 );
	int mhz = 0; cudaDeviceGetAttribute( &mhz, cudaDevAttrClockRate, 0 );
	const double winst = (double)grid * 8 /* warps */ * ITERS * U * per_iter;
	printf( "%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, winst / (ms * 1e-3) / (mhz * 1e3) / 148, mhz / 1000 );
}
int main()
{
	float* d; cudaMalloc( &d, 148 * 8 * 256 * 4 );
	run<0>( "FFMA", 1, d ); run<1>( "FFMA2", 1, d ); run<2>( "FADD", 1, d ); run<3>( "FADD2", 1, d ); run<4>( "PRMT+LOP3", 2, d );
	run<5>( "FMNMX3", 1, d ); run<6>( "I2F.U8 (+shift)", 2, d ); run<7>( "FFMA + PRMT", 2, d ); run<8>( "FFMA2 + PRMT + LOP3", 3, d );
	printf( "%s\n", cudaGetErrorString( cudaDeviceSynchronize() ) );
	return 0;
}
