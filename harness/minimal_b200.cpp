// harness/minimal_b200.cpp - smallest C++ use of the shim (the role tiny_bvh_minimal_gpu.cpp plays in the reference):
// random triangles -> BVH::Build on the GPU -> a batch of rays -> nearest hits.  No reference header needed.
//   g++ -O2 -Iinclude harness/minimal_b200.cpp -Ltinybvh_b200 -ltinybvh_b200 -Wl,-rpath,$PWD/tinybvh_b200 -o minimal_b200
#include "tinybvh_b200.hpp"
#include <vector>

struct Vec4 { float x, y, z, w; };
static uint32_t seed = 0x12345678;
static float rnd() { seed ^= seed << 13, seed ^= seed >> 17, seed ^= seed << 5; return seed * 2.3283064365387e-10f; }

int main()
{
	const int N = 8192, R = 1024;
	std::vector<Vec4> tris( N * 3 );
	for (int i = 0; i < N; i++)
	{
		const float x = rnd() * 4, y = rnd() * 4, z = rnd() * 4;
		for (int v = 0; v < 3; v++) tris[i * 3 + v] = { x + rnd() * 0.1f, y + rnd() * 0.1f, z + rnd() * 0.1f, 0 };
	}
	tinybvh_b200::BVH bvh;
	bvh.Build( tris.data(), N );
	tinybvh_b200::Ray* rays = (tinybvh_b200::Ray*)tinybvh_b200::malloc_pinned( R * sizeof( tinybvh_b200::Ray ) );
	for (int i = 0; i < R; i++)
	{
		const float O[3] = { 2, 2, -3 }, D[3] = { (i % 32) / 32.0f - 0.5f, (i / 32) / 32.0f - 0.5f, 1 };
		rays[i] = tinybvh_b200::Ray( O, D );
	}
	bvh.Intersect( rays, R );
	int hits = 0;
	double sum = 0;
	for (int i = 0; i < R; i++) if (rays[i].t < 1e30f) hits++, sum += rays[i].t;
	printf( "minimal_b200: %i tris, %u nodes, build %.3f ms; %i of %i rays hit, mean t %.4f, ray 528: t=%f prim=%u\n",
		N, bvh.usedNodes, bvh.buildMs, hits, R, hits ? sum / hits : 0.0, rays[528].t, rays[528].prim );
	tinybvh_b200::free_pinned( rays );
	return hits > 0 ? 0 : 1;
}
