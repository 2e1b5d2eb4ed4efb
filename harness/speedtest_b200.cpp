// harness/speedtest_b200.cpp - the GPU section of the reference's benchmark, re-hosted on the B200 engine.
//
// This is the drop-in demonstration for SURVEY.md 8(b): the host logic is the speedtest's (tiny_bvh_speedtest.cpp): load
// ./testdata-style .bin, the three speedtest cameras, 4x4-tile primary rays (:497-551), shadow rays (:844-865), the
// CPU reference distances (:560-574), ValidateTraceResult's tolerances (:338-381) - and where the speedtest calls
// tinyocl (Kernel / Buffer / Run, :1092-1241) it calls the tinybvh_b200 shim instead.  It #includes the UNMODIFIED
// reference header for the host types (tinybvh::Ray, bvhvec4) and the CPU reference, so it is compiled where
// /root/reference exists (oracle/Makefile target `speedtest`, output oracle/_ref/speedtest_b200) and only the binary
// travels.  TEST / DEMONSTRATION code: it links the reference, so it is not part of the product library.
//
// usage: speedtest_b200 <scene.bin> [width height]       (defaults 800 x 600)
#define TINYBVH_IMPLEMENTATION
#include "tiny_bvh.h"
#include "tinybvh_b200.hpp"
#include <atomic>
#include <chrono>
#include <fstream>
#include <thread>
#include <vector>

using namespace tinybvh;

struct Timer
{
	std::chrono::high_resolution_clock::time_point start = std::chrono::high_resolution_clock::now();
	float elapsed() const { return std::chrono::duration<float>( std::chrono::high_resolution_clock::now() - start ).count(); }
	void reset() { start = std::chrono::high_resolution_clock::now(); }
};

static int W = 800, H = 600;
static bvhvec4* triangles = 0;
static uint32_t verts = 0;

// CPU reference over a batch with all host threads: 10,000-ray batches off an atomic counter (:387-401)
template <class F> static void parallel_batches( size_t n, F f )
{
	const unsigned threads = std::max( 1u, std::thread::hardware_concurrency() );
	const size_t B = 10000, batches = (n + B - 1) / B;
	std::atomic<size_t> next( 0 );
	std::vector<std::thread> pool;
	for (unsigned t = 0; t < threads; t++) pool.emplace_back( [&]() { for (;;) { size_t b = next++; if (b >= batches) break; f( b * B, std::min( n, b * B + B ) ); } } );
	for (auto& t : pool) t.join();
}

// the speedtest's statistical validator (:338-381): sum of t over every 4th ray within 1 %, sums of u, v within 5 %
static bool validate_like_speedtest( const Ray* got, const Ray* want, size_t n )
{
	double st = 0, su = 0, sv = 0, rt = 0, ru = 0, rv = 0;
	for (size_t i = 0; i < n; i += 4)
	{
		st += got[i].hit.t == 1e30f ? 100 : got[i].hit.t, rt += want[i].hit.t == 1e30f ? 100 : want[i].hit.t;
		if (want[i].hit.t < 100) su += got[i].hit.u, sv += got[i].hit.v, ru += want[i].hit.u, rv += want[i].hit.v;
	}
	const bool ok = fabs( st - rt ) <= 0.01 * fabs( rt ) && fabs( su - ru ) <= 0.05 * fabs( ru ) + 1e-6 && fabs( sv - rv ) <= 0.05 * fabs( rv ) + 1e-6;
	if (!ok) printf( "!! Validation failed: sum t %f vs %f, u %f vs %f, v %f vs %f\n", st, rt, su, ru, sv, rv );
	return ok;
}

int main( int argc, char** argv )
{
	if (argc < 2) { printf( "usage: %s scene.bin [width height]\n", argv[0] ); return 2; }
	if (argc >= 4) W = atoi( argv[2] ), H = atoi( argv[3] );
	std::fstream s{ argv[1], s.binary | s.in };
	if (!s) { printf( "cannot open %s\n", argv[1] ); return 2; }
	s.read( (char*)&verts, 4 );
	printf( "Loading triangle data (%u tris).\n", verts );
	verts *= 3, triangles = (bvhvec4*)malloc64( verts * sizeof( bvhvec4 ) );
	s.read( (char*)triangles, verts * 16 );

	// cameras and primary rays exactly as the speedtest sets them up (:497-551)
	bvhvec3 eyes[3] = { bvhvec3( -15.24f, 21.5f, 2.54f ), bvhvec3( -34, 5, 11.26f ), bvhvec3( -1.3, 4.96, 12.28 ) };
	bvhvec3 views[3] = { tinybvh_normalize( bvhvec3( 0.826f, -0.438f, -0.356f ) ), tinybvh_normalize( bvhvec3( 0.9427, 0.0292, -0.3324 ) ), tinybvh_normalize( bvhvec3( -0.9886, 0.0507, -0.1419 ) ) };
	const size_t Nfull = (size_t)W * H * 16;
	Ray* fullBatch[3], * refBatch[3];
	for (int i = 0; i < 3; i++)
	{
		const bvhvec3 eye = eyes[i], view = views[i];
		const bvhvec3 right = tinybvh_normalize( tinybvh_cross( bvhvec3( 0, 1, 0 ), view ) ), up = 0.8f * tinybvh_cross( view, right ), C = eye + 2 * view;
		const bvhvec3 p1 = C - right + up, p2 = C + right + up, p3 = C - right - up;
		fullBatch[i] = (Ray*)tinybvh_b200::malloc_pinned( Nfull * sizeof( Ray ) ); // was tinybvh::malloc64 (:520)
		refBatch[i] = (Ray*)malloc64( Nfull * sizeof( Ray ) );
		size_t n = 0;
		for (int ty = 0; ty < H / 4; ty++) for (int tx = 0; tx < W / 4; tx++) for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
			for (int sm = 0; sm < 16; sm++)
			{
				const float u = (float)((tx * 4 + x) * 4 + (sm & 3)) / (W * 4), v = (float)((ty * 4 + y) * 4 + (sm >> 2)) / (H * 4);
				const bvhvec3 P = p1 + u * (p2 - p1) + v * (p3 - p1);
				fullBatch[i][n++] = Ray( eye, tinybvh_normalize( P - eye ) );
			}
		memcpy( refBatch[i], fullBatch[i], Nfull * sizeof( Ray ) );
	}

	// ---- CPU reference: BVH::Build + BVH::Intersect, all host threads (:1076-1090)
	Timer t;
	BVH ref_bvh;
	ref_bvh.Build( triangles, verts / 3 );
	const float refBuild = t.elapsed();
	printf( "reference BVH::Build          : %7.2f ms, %u nodes, SAH %.2f\n", refBuild * 1000, ref_bvh.usedNodes, ref_bvh.SAHCost() );
	t.reset();
	for (int i = 0; i < 3; i++) parallel_batches( Nfull, [&]( size_t a, size_t b ) { for (size_t k = a; k < b; k++) ref_bvh.Intersect( refBatch[i][k] ); } );
	const float refTrace = t.elapsed() / 3;
	printf( "reference BVH::Intersect      : %7.2f ms per view, %.1f Mrays/s on %u threads\n", refTrace * 1000, Nfull / refTrace / 1e6, std::thread::hardware_concurrency() );

	// ---- GPU section (replaces :1092-1241): build on the GPU, trace the three views, copy hits back into the Ray array
	tinybvh_b200::BVH gpu_bvh;
	gpu_bvh.Build( triangles, verts / 3 ); // warm-up (first-launch cost), then the timed build
	gpu_bvh.Build( triangles, verts / 3 );
	printf( "tinybvh_b200 BVH::Build       : %7.3f ms (device), %u nodes  -> %.1f Mtris/s\n", gpu_bvh.buildMs, gpu_bvh.usedNodes, verts / 3 / gpu_bvh.buildMs / 1e3 );
	// a GPU-built tree is usable by the reference's own host code: download into a reference BVH and ask it for SAHCost
	{
		BVH host;
		host.bvhNode = (BVH::BVHNode*)host.AlignedAlloc( (size_t)gpu_bvh.usedNodes * 32 ), host.primIdx = (uint32_t*)host.AlignedAlloc( (size_t)gpu_bvh.idxCount * 4 );
		gpu_bvh.Download( host.bvhNode, host.primIdx );
		host.usedNodes = host.allocatedNodes = gpu_bvh.usedNodes, host.idxCount = gpu_bvh.idxCount, host.triCount = verts / 3;
		host.verts = bvhvec4slice( triangles, verts, sizeof( bvhvec4 ) );
		const bool same = gpu_bvh.usedNodes == ref_bvh.usedNodes && !ref_bvh.threadedBuild && memcmp( host.bvhNode, ref_bvh.bvhNode, (size_t)gpu_bvh.usedNodes * 32 ) == 0;
		printf( "  downloaded tree: SAH %.2f (reference %.2f)%s\n", host.SAHCost(), ref_bvh.SAHCost(), same ? ", node array identical to BVH::Build" : "" );
	}
	gpu_bvh.Intersect( fullBatch[0], Nfull ); // warm-up pass (the speedtest drops pass 0, :1222-1231)
	for (size_t k = 0; k < Nfull; k++) fullBatch[0][k].hit.t = 1e30f;
	t.reset();
	for (int i = 0; i < 3; i++) gpu_bvh.Intersect( fullBatch[i], Nfull );
	const float gpuTrace = t.elapsed() / 3;
	printf( "tinybvh_b200 BVH::Intersect   : %7.2f ms per view incl. PCIe both ways, %.1f Mrays/s\n", gpuTrace * 1000, Nfull / gpuTrace / 1e6 );

	// ---- validation: the speedtest's tolerance, then the exact per-ray comparison this repo holds itself to
	bool ok = true;
	size_t primDiff = 0, tDiff = 0;
	for (int i = 0; i < 3; i++)
	{
		ok &= validate_like_speedtest( fullBatch[i], refBatch[i], Nfull );
		for (size_t k = 0; k < Nfull; k++)
		{
			const bool hit = refBatch[i][k].hit.t < 1e30f;
			primDiff += hit && fullBatch[i][k].hit.prim != refBatch[i][k].hit.prim;
			tDiff += memcmp( &fullBatch[i][k].hit.t, &refBatch[i][k].hit.t, 4 ) != 0;
		}
	}
	printf( "validation: speedtest tolerance %s; exact: %zu prim mismatches, %zu t-bit mismatches over %zu rays\n", ok ? "passed" : "FAILED", primDiff, tDiff, 3 * Nfull );

	// ---- SBVH (the speedtest's BuildHQ line, :678): reference BVH::BuildHQ on the host, the same builder on the GPU
	bool hqSame = false;
	{
		BVH ref_hq;
		t.reset();
		ref_hq.BuildHQ( triangles, verts / 3 );
		const float refHQ = t.elapsed();
		printf( "reference BVH::BuildHQ        : %7.2f ms, %u nodes, SAH %.2f\n", refHQ * 1000, ref_hq.usedNodes, ref_hq.SAHCost() );
		tinybvh_b200::BVH gpu_hq;
		gpu_hq.BuildHQ( triangles, verts / 3 );
		gpu_hq.BuildHQ( triangles, verts / 3 );
		printf( "tinybvh_b200 BVH::BuildHQ     : %7.3f ms (device), %u nodes, idxCount %u  -> %.1fx\n", gpu_hq.buildMs, gpu_hq.usedNodes, gpu_hq.idxCount, refHQ * 1000 / gpu_hq.buildMs );
		std::vector<uint8_t> nodes( (size_t)gpu_hq.usedNodes * 32 );
		std::vector<uint32_t> idx( gpu_hq.idxCount );
		gpu_hq.Download( nodes.data(), idx.data() );
		uint32_t refs = 0;
		for (uint32_t n = 0; n < ref_hq.usedNodes; n++) if (n != 1) refs += ref_hq.bvhNode[n].triCount;
		hqSame = gpu_hq.usedNodes == ref_hq.usedNodes && gpu_hq.idxCount == ref_hq.idxCount && memcmp( nodes.data(), ref_hq.bvhNode, nodes.size() ) == 0 &&
			memcmp( idx.data(), ref_hq.primIdx, (size_t)refs * 4 ) == 0;
		printf( "  SBVH node array and the %u referenced primIdx entries %s BVH::BuildHQ's\n", refs, hqSame ? "are identical to" : "DIFFER from" );
	}

	// ---- animation: move the vertices in place, Refit both trees (tiny_bvh_anim.cpp's pattern; BVH::Refit :3055)
	bool refitSame = false;
	{
		BVH ref_anim;
		ref_anim.threadedBuild = false;
		ref_anim.Build( triangles, verts / 3 );
		tinybvh_b200::BVH gpu_anim;
		gpu_anim.Build( triangles, verts / 3 );
		for (uint32_t i = 0; i < verts; i++) triangles[i].y += 0.05f * sinf( triangles[i].x * 0.37f ); // the caller edits its own array
		t.reset();
		ref_anim.Refit();
		const float refRefit = t.elapsed();
		gpu_anim.Refit();
		std::vector<uint8_t> nodes( (size_t)gpu_anim.usedNodes * 32 );
		std::vector<uint32_t> idx( gpu_anim.idxCount );
		gpu_anim.Download( nodes.data(), idx.data() );
		refitSame = gpu_anim.usedNodes == ref_anim.usedNodes && memcmp( nodes.data(), ref_anim.bvhNode, nodes.size() ) == 0;
		printf( "BVH::Refit: reference %.2f ms, tinybvh_b200 %.3f ms (device); refitted node array %s\n", refRefit * 1000, gpu_anim.buildMs, refitSame ? "identical" : "DIFFERS" );
	}

	// ---- TLAS / BLAS (tiny_bvh_gpu2.cpp's pattern): 27 instances of the scene, reference IntersectTLAS vs the two-level kernel
	bool tlasSame = false;
	{
		BVH ref_blas;
		ref_blas.Build( triangles, verts / 3 );
		tinybvh_b200::BVH gpu_blas;
		gpu_blas.Build( triangles, verts / 3 );
		const int N = 27;
		BLASInstance* inst = (BLASInstance*)malloc64( N * sizeof( BLASInstance ) );
		const float span = 2.2f * tinybvh_max( ref_blas.aabbMax.x - ref_blas.aabbMin.x, ref_blas.aabbMax.z - ref_blas.aabbMin.z );
		for (int i = 0; i < N; i++)
		{
			inst[i] = BLASInstance( 0 );
			const float a = 0.4f * i, sc = 0.5f + 0.04f * i;
			float* T = (float*)&inst[i].transform;
			T[0] = cosf( a ) * sc, T[2] = sinf( a ) * sc, T[5] = sc, T[8] = -sinf( a ) * sc, T[10] = cosf( a ) * sc, T[15] = 1;
			T[1] = T[4] = T[6] = T[9] = T[12] = T[13] = T[14] = 0;
			T[3] = (i % 3 - 1) * span, T[7] = (i / 9 - 1) * span * 0.6f, T[11] = ((i / 3) % 3 - 1) * span;
		}
		BLASInstance* inst2 = (BLASInstance*)malloc64( N * sizeof( BLASInstance ) );
		memcpy( inst2, inst, N * sizeof( BLASInstance ) );
		BVHBase* blasList[1] = { &ref_blas };
		BVH ref_tlas;
		ref_tlas.Build( inst, N, blasList, 1 ); // Update()s the instances: inverse transforms + world boxes
		tinybvh_b200::BVHBase* gpuList[1] = { &gpu_blas };
		tinybvh_b200::BVH gpu_tlas;
		gpu_tlas.Build( inst2, N, gpuList, 1 ); // the shim Update()s its copy the same way
		const bool updSame = memcmp( inst, inst2, N * sizeof( BLASInstance ) ) == 0;
		printf( "BLASInstance::Update: %d records %s the reference's\n", N, updSame ? "identical to" : "DIFFER from" );
		const size_t M = (size_t)W * H;
		Ray* a = (Ray*)malloc64( M * sizeof( Ray ) ), * b = (Ray*)tinybvh_b200::malloc_pinned( M * sizeof( Ray ) ); // rays that cross PCIe: page-locked
		const bvhvec3 eye( span * 0.2f, span * 1.1f, -span * 2.6f ), view = tinybvh_normalize( bvhvec3( -0.05f, -0.35f, 1 ) );
		const bvhvec3 right = tinybvh_normalize( tinybvh_cross( bvhvec3( 0, 1, 0 ), view ) ), up = 0.75f * tinybvh_cross( view, right ), C = eye + 1.2f * view;
		for (int y = 0; y < H; y++) for (int x = 0; x < W; x++)
		{
			const bvhvec3 P = C + ((x + 0.5f) / W * 2 - 1) * right + (1 - (y + 0.5f) / H * 2) * up;
			a[(size_t)y * W + x] = Ray( eye, tinybvh_normalize( P - eye ) );
		}
		memcpy( b, a, M * sizeof( Ray ) );
		t.reset();
		parallel_batches( M, [&]( size_t s0, size_t s1 ) { for (size_t k = s0; k < s1; k++) ref_tlas.Intersect( a[k] ); } );
		const float refT = t.elapsed();
		gpu_tlas.Intersect( b, M ); // warm-up, then timed on fresh rays
		for (size_t k = 0; k < M; k++) b[k].hit = Ray( eye, view ).hit;
		t.reset();
		gpu_tlas.Intersect( b, M );
		const float gpuT = t.elapsed();
		size_t diff = 0, hits = 0;
		for (size_t k = 0; k < M; k++)
		{
			hits += a[k].hit.t < 1e30f;
			diff += memcmp( &a[k].hit.t, &b[k].hit.t, 16 ) != 0 || (a[k].hit.t < 1e30f && a[k].hit.inst != b[k].hit.inst);
		}
		tlasSame = diff == 0 && gpu_tlas.usedNodes == ref_tlas.usedNodes && updSame;
		printf( "TLAS of %d instances: reference IntersectTLAS %.2f ms (%u threads), tinybvh_b200 %.2f ms incl. PCIe; %zu of %zu rays hit, %zu differ (t,u,v,prim,inst)\n",
			N, refT * 1000, std::thread::hardware_concurrency(), gpuT * 1000, hits, M, diff );
		// the same instances over a BVH8_CWBVH BLAS (tiny_bvh_gpu2.cpp's arrangement; the reference's CPU IntersectTLAS does not take one):
		// reported, not part of the verdict - two layouts of the same triangles agree up to exact-distance ties
		{
			tinybvh_b200::BVH8_CWBVH gpu_wide;
			gpu_wide.Build( triangles, verts / 3 );
			tinybvh_b200::BVHBase* wideList[1] = { &gpu_wide };
			tinybvh_b200::BVH gpu_tlas_wide;
			gpu_tlas_wide.Build( inst2, N, wideList, 1 );
			for (size_t k = 0; k < M; k++) b[k].hit = Ray( eye, view ).hit;
			gpu_tlas_wide.Intersect( b, M );
			for (size_t k = 0; k < M; k++) b[k].hit = Ray( eye, view ).hit;
			t.reset();
			gpu_tlas_wide.Intersect( b, M );
			const float wideT = t.elapsed();
			size_t wdiff = 0;
			for (size_t k = 0; k < M; k++) wdiff += memcmp( &a[k].hit.t, &b[k].hit.t, 16 ) != 0 || (a[k].hit.t < 1e30f && a[k].hit.inst != b[k].hit.inst);
			printf( "TLAS over a BVH8_CWBVH BLAS (layout %d): tinybvh_b200 %.2f ms incl. PCIe; %zu of %zu rays differ from the reference's IntersectTLAS over the BVH BLAS\n",
				gpu_tlas_wide.Layout(), wideT * 1000, wdiff, M );
		}
	}
	return ok && primDiff == 0 && tDiff == 0 && hqSame && refitSame && tlasSame ? 0 : 1;
}
