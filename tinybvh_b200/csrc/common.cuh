// tinybvh_b200/csrc/common.cuh - shared device/host definitions of the sm_100a engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tinybvh_b200.h"

#define BVH_FAR 1e30f
#define TBVH_STACK 64          // traversal stack entries per ray (reference: 256 closest / 64 any-hit, tiny_bvh.h:3249,:3409)

// ---- error plumbing -------------------------------------------------------------------------------------
void tbvh_set_error( const char* fmt, ... );
extern unsigned long long g_tbvh_launches;
#define CUDA_TRY( x ) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { tbvh_set_error( "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString( e_ ) ); return TBVH_E_CUDA; } } while (0)
#define LAUNCHED() do { g_tbvh_launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { tbvh_set_error( "%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString( e_ ) ); return TBVH_E_CUDA; } } while (0)

// ---- handles ----------------------------------------------------------------------------------------------
struct tbvh_ctx_t
{
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = 0;         // engine stream (builds, uploads, host-path copies)
	cudaStream_t copy_streams[3] = { 0, 0, 0 }; // host-path pipeline
	void* d_stage[3] = { 0, 0, 0 };  // staging ray chunks for the host path
	void* d_stage_bits[3] = { 0, 0, 0 };
	size_t stage_rays = 0;
	// host-path tuning (env TBVH_HOST_PATH = copy2d | zerocopy, TBVH_H2D_SPLIT = 1..4): the inbound 2D copy of a chunk
	// can be split over several streams so more than one copy engine works on it
	int host_path = 0;               // 0 = copy engine (cudaMemcpy2DAsync), 1 = copy kernels through the pinned mapping
	int h2d_split = 1;
	int trace_variant = 3;           // BVH2 traversal kernel: 0 generic, 3 octant switch, 4 persistent warps (see trace_bvh2.cu)
	int small_mode = 0;              // warp-subtree kernel: bit 0 = fragments staged in shared memory, bit 1 = aggregated bin updates
	int inst_idx_bits = 32;          // the host program's INST_IDX_BITS (tiny_bvh.h:118): 32 = TLAS hits store hit.inst, 4..31 = top bits of hit.prim
	int hq_small = 16;               // BuildHQ: nodes of at most this many fragments go to the warp-per-subtree kernel (<= 256)
	int hq_cluster = 16;             // BuildHQ: largest thread-block cluster a node of the level phase may get (1..16)
	int small_t = 128;               // builder: subtrees of at most this many primitives go to the warp kernel (<= 256)
	int d2h_mode = 0;                // hits back to the host: 0 = 2D copy of 16-byte rows, 1 = 2D copy of the whole 64-byte rows,
	                                 // 2 = packed 1D copy to pinned staging + multi-threaded host scatter, 3 = scatter kernel (zero copy)
	void* d_hits_pack[3] = { 0, 0, 0 };
	// host_path 2: rays are packed by host threads into pinned staging (48 or 32 bytes per ray) and cross PCIe as ONE contiguous copy
	void* h_pack[3] = { 0, 0, 0 }; void* d_pack[3] = { 0, 0, 0 };
	cudaEvent_t ev_pack[3] = { 0, 0, 0 };
	void* h_hits = 0; size_t h_hits_rays = 0; // pinned staging for d2h_mode 2
	cudaStream_t aux_streams[4] = { 0, 0, 0, 0 };
	cudaEvent_t ev_done[3] = { 0, 0, 0 };
	cudaEvent_t ev_part[3][4] = {};
};

struct tbvh_bvh_t
{
	tbvh_ctx ctx = 0;
	tbvh_info info = {};
	// geometry (engine-owned copy, float4 per vertex)
	float4* d_verts = 0;
	// LAYOUT_BVH: reference node array; children of an interior node are the 64-byte pair at nodes[leftFirst]
	float4* d_nodes = 0;       // 2 float4 per node
	uint32_t* d_prim_idx = 0;
	// traversal view of the BVH2: d_trav aliases d_nodes (LAYOUT_BVH) or is the pair array derived from a BVH_GPU upload
	float4* d_trav = 0;
	uint32_t root_ref = 0, root_count = 0; // the root as a child record: count==0 -> pair index, else leaf range
	// leaf-ordered triangle records for BVH2 traversal: 3 float4 per prim reference
	//   [0] = (v0.xyz, as_float(primIdx))  [1] = e1 = v1-v0  [2] = e2 = v2-v0
	float4* d_leaf_tris = 0;
	// LAYOUT_BVH_GPU mirror (only materialised on upload / convert / download)
	float4* d_nodes_gpu = 0;   // 4 float4 per node
	// LAYOUT_CWBVH
	float4* d_cw_nodes = 0;    // 5 float4 per node
	float4* d_cw_tris = 0;     // 3 float4 per triangle
	// TLAS (BVH::Build( BLASInstance*, instCount, BVHBase**, blasCount ) :2221): nodes / primIdx over instance boxes + device tables
	float4* d_aabbs = 0;       // instance boxes the TLAS was built over (2 float4 per instance)
	void* d_inst = 0;          // TlasInst records (inverse transform, BLAS number, mask)
	void* d_blas = 0;          // BlasRef records (traversal arrays of every BLAS)
	uint32_t inst_count = 0, blas_count = 0;
	bool refittable = true;    // BVHBase::refittable (:811): false after BuildHQ ("can't refit an SBVH", :3027)
	// statistics
	int stats = 0;
	unsigned long long* d_stats = 0; // [0]=steps [1]=tris
};

// ---- device math in the oracle's exact operation order (oracle/tbvh_oracle.c header lists the pairing) ----
// Every fused pair is spelled __fmaf_rn, every unfused product / sum an _rn intrinsic, so nvcc's own
// contraction (-fmad) cannot change the rounding.

// MOLLER_TRUMBORE_TEST tiny_bvh.h:1644-1656 with e1,e2 precomputed (identical bits: v1-v0 is exact-rounded once).
// Returns true when the triangle is accepted for [0, tmax]; writes t,u,v.
__device__ __forceinline__ bool mt_test( const float ox, const float oy, const float oz, const float dx, const float dy, const float dz,
	const float4 v0, const float4 e1, const float4 e2, const float tmax, float& t, float& u, float& v )
{
	const float hx = __fmaf_rn( dy, e2.z, -__fmul_rn( dz, e2.y ) );
	const float hy = __fmaf_rn( dz, e2.x, -__fmul_rn( dx, e2.z ) );
	const float hz = __fmaf_rn( dx, e2.y, -__fmul_rn( dy, e2.x ) );
	const float a = __fmaf_rn( e1.z, hz, __fmaf_rn( e1.x, hx, __fmul_rn( e1.y, hy ) ) );
	if (fabsf( a ) < 0.000001f) return false;
	const float f = __fdiv_rn( 1.0f, a );
	const float sx = __fsub_rn( ox, v0.x ), sy = __fsub_rn( oy, v0.y ), sz = __fsub_rn( oz, v0.z );
	u = __fmul_rn( f, __fmaf_rn( hz, sz, __fmaf_rn( hx, sx, __fmul_rn( hy, sy ) ) ) );
	const float qx = __fmaf_rn( -e1.y, sz, __fmul_rn( e1.z, sy ) );
	const float qy = __fmaf_rn( -e1.z, sx, __fmul_rn( e1.x, sz ) );
	const float qz = __fmaf_rn( -e1.x, sy, __fmul_rn( e1.y, sx ) );
	v = __fmul_rn( f, __fmaf_rn( dz, qz, __fmaf_rn( dy, qy, __fmul_rn( dx, qx ) ) ) );
	if (u < 0 || v < 0 || __fadd_rn( u, v ) > 1) return false;
	t = __fmul_rn( f, __fmaf_rn( e2.z, qz, __fmaf_rn( e2.x, qx, __fmul_rn( e2.y, qy ) ) ) );
	return !(t < 0 || t > tmax);
}

// order-preserving float <-> uint key for atomicMin/Max on floats
__device__ __forceinline__ uint32_t f2key( float f ) { uint32_t u = __float_as_uint( f ); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f( uint32_t k ) { return __uint_as_float( (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k ); }

// ---- internal entry points (one per .cu) -------------------------------------------------------------------
int bvh2_trace_launch( tbvh_bvh b, const void* d_rays, uint32_t stride, void* d_hits, uint32_t hit_stride, uint32_t* d_bits, uint64_t n, bool anyhit, cudaStream_t s );
int cwbvh_trace_launch( tbvh_bvh b, const void* d_rays, uint32_t stride, void* d_hits, uint32_t hit_stride, uint32_t* d_bits, uint64_t n, bool anyhit, cudaStream_t s );
int build_sah_launch( tbvh_bvh b, float c_trav, float c_int, int flavour );
int build_hq_launch( tbvh_bvh b, float c_trav, float c_int );
int refit_launch( tbvh_bvh b, cudaStream_t s );
int tlas_trace_launch( tbvh_bvh b, const void* d_rays, uint32_t stride, uint32_t* d_bits, uint64_t n, bool anyhit, cudaStream_t s );
struct TlasInst { float inv[16]; uint32_t blasIdx, mask, pad0, pad1; };                                  // 80 bytes
struct BlasRef { const float4* trav; const float4* tris; uint32_t root_ref, root_count, pad0, pad1; };   // 32 bytes
int make_leaf_tris( tbvh_bvh b, cudaStream_t s );
int bvh_gpu_to_bvh( tbvh_bvh b, uint32_t used_nodes_gpu, cudaStream_t s );
int bvh_to_bvh_gpu( tbvh_bvh b, cudaStream_t s );
int bvh_to_cwbvh( tbvh_bvh b, cudaStream_t s );
// exclusive scan of in[0..n) into out[0..n] (out[n] = total); tile_sum needs n/2048 + 2 words (build_sah.cu)
int exclusive_scan( const uint32_t* in, uint32_t* out, uint32_t* tile_sum, uint32_t n, cudaStream_t s );
