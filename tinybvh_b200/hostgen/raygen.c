/* tinybvh_b200/hostgen/raygen.c - multi-threaded generators for the synthetic ray sets of SURVEY.md 8(d) (workload side of the
 * bench and the tests; not part of the engine).  Every function is the float32 arithmetic of tinybvh_b200/rays.py, operation for
 * operation (compile with -ffp-contract=off), which in turn restates Ray::Ray (tiny_bvh.h:695-701), tinybvh_normalize (:508-512),
 * tinybvh_safercp (:442) and the ray loops of tiny_bvh_speedtest.cpp (:497-551 camera tiles, :853-865 shadow rays, :564-587 bounce
 * rays).  numpy builds 16.8 M of these records in ~30 s on one core; the bench needs 2 x 67.1 M.
 *
 * Records are the reference's 128-byte host Ray: O @0, mask @12, D @16, instIdx @28, rD @32, t @48, u @52, v @56, prim @60. */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>

/* launchers such as torchrun export OMP_NUM_THREADS=1; the generators are told explicitly how many threads they may use */
void tbvh_gen_set_threads( const int n ) { if (n > 0) omp_set_num_threads( n ); }

#define BVH_FAR 1e30f

static inline float safercp( const float x ) { return (x > 1e-12f || x < -1e-12f) ? 1.0f / x : (x >= 0 ? BVH_FAR : -BVH_FAR); }

static inline void put_ray( char* rec, const float* O, const float* D, const float tmax )
{
	float* f = (float*)rec;
	memset( rec, 0, 128 );
	f[0] = O[0], f[1] = O[1], f[2] = O[2];
	((uint32_t*)rec)[3] = 0xFFFFu;
	f[4] = D[0], f[5] = D[1], f[6] = D[2];
	f[8] = safercp( D[0] ), f[9] = safercp( D[1] ), f[10] = safercp( D[2] );
	f[12] = tmax;
}

static inline void normalize3( float* v )
{
	const float l = sqrtf( v[0] * v[0] + v[1] * v[1] + v[2] * v[2] );
	const float rl = l == 0 ? 0.0f : 1.0f / l;
	v[0] = v[0] * rl, v[1] = v[1] * rl, v[2] = v[2] * rl;
}

/* rays [first, first+count) of the width x height x spp camera set: 4x4-pixel tiles, sample s of a pixel on the 4x4 sub-grid.
 * p1 = top-left corner of the view plane, d21 = p2 - p1, d31 = p3 - p1 (computed by the caller as rays.py does). */
void tbvh_gen_primary( char* out, const uint64_t first, const uint64_t count, const float* eye, const float* p1, const float* d21, const float* d31,
	const uint32_t width, const uint32_t height, const uint32_t spp )
{
	const uint32_t tiles_x = width / 4, step = 16 / spp;
	const float fw = (float)(width * 4), fh = (float)(height * 4);
	(void)height;
	#pragma omp parallel for schedule( static )
	for (int64_t k = 0; k < (int64_t)count; k++)
	{
		const uint64_t i = first + (uint64_t)k;
		const uint32_t si = (uint32_t)(i % spp);
		uint64_t r = i / spp;
		const uint32_t x = (uint32_t)(r & 3), y = (uint32_t)((r >> 2) & 3);
		r >>= 4;
		const uint32_t tx = (uint32_t)(r % tiles_x), ty = (uint32_t)(r / tiles_x);
		const uint32_t px = tx * 4 + x, py = ty * 4 + y, s = si * step;
		const float u = (float)(px * 4 + (s & 3)) / fw, v = (float)(py * 4 + (s >> 2)) / fh;
		float D[3];
		for (int a = 0; a < 3; a++)
		{
			const float P = (p1[a] + u * d21[a]) + v * d31[a];
			D[a] = P - eye[a];
		}
		normalize3( D );
		normalize3( D ); /* the caller normalises, Ray::Ray normalises again (tiny_bvh.h:698) - rays.py make_rays( .., normalized = False ) */
		put_ray( out + (uint64_t)k * 128, eye, D, BVH_FAR );
	}
}

/* shadow rays from traced primaries (tiny_bvh_speedtest.cpp:853-865): `prim` are 128-byte records; when `hits` is not NULL the hit
 * distance of ray i is hits[4*i] (a packed t,u,v,prim array) instead of the record's own hit.t. */
void tbvh_gen_shadow( char* out, const char* prim, const float* hits, const uint64_t count, const float* light, const float eps )
{
	#pragma omp parallel for schedule( static )
	for (int64_t k = 0; k < (int64_t)count; k++)
	{
		const float* p = (const float*)(prim + (uint64_t)k * 128);
		const float th = hits ? hits[(uint64_t)k * 4] : p[12];
		const float t = th < 1000.0f ? th : 1000.0f; /* np.minimum( 1000, t ) */
		float I[3], toL[3], D[3], O[3];
		for (int a = 0; a < 3; a++) I[a] = p[a] + t * p[4 + a], toL[a] = light[a] - I[a], D[a] = toL[a];
		normalize3( D );
		const float dist = sqrtf( toL[0] * toL[0] + toL[1] * toL[1] + toL[2] * toL[2] );
		for (int a = 0; a < 3; a++) O[a] = I[a] + D[a] * eps;
		put_ray( out + (uint64_t)k * 128, O, D, dist - eps );
	}
}

static inline uint32_t wang( uint32_t seed )
{
	uint32_t s = (seed ^ 61u) ^ (seed >> 16);
	s *= 9u, s ^= s >> 4, s *= 0x27D4EB2Du, s ^= s >> 15;
	return s;
}
static inline uint32_t xorshift( uint32_t s ) { s ^= s << 13, s ^= s >> 17, s ^= s << 5; return s; }

/* diffuse-bounce rays (tiny_bvh_speedtest.cpp:564-587 with rand() replaced by a per-ray xorshift32):
 * ray `first + k` of the whole set is seeded with wang( (first + k + seed) | 1 ).  verts = float4 x 3 per triangle. */
void tbvh_gen_diffuse( char* out, const char* prim, const float* hits, const uint64_t first, const uint64_t count, const float* verts, const uint32_t seed )
{
	#pragma omp parallel for schedule( static )
	for (int64_t k = 0; k < (int64_t)count; k++)
	{
		const float* p = (const float*)(prim + (uint64_t)k * 128);
		const float t = hits ? hits[(uint64_t)k * 4] : p[12];
		const uint32_t pr = hits ? ((const uint32_t*)hits)[(uint64_t)k * 4 + 3] : ((const uint32_t*)p)[15];
		uint32_t s = wang( ((uint32_t)(first + (uint64_t)k) + seed) | 1u );
		float R[3];
		for (int a = 0; a < 3; a++) { s = xorshift( s ); R[a] = (float)((double)s * 2.3283064365387e-10) - 0.5f; }
		normalize3( R );
		const float* O = p, * D = p + 4;
		const int hit = t < 100.0f;
		float I[3], N[3];
		for (int a = 0; a < 3; a++) I[a] = hit ? O[a] + t * D[a] : O[a] + 20.0f * D[a];
		const float* v0 = verts + (uint64_t)(hit ? pr : 0) * 12, * v1 = v0 + 4, * v2 = v0 + 8;
		const float e1[3] = { v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2] }, e2[3] = { v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2] };
		N[0] = e1[1] * e2[2] - e1[2] * e2[1], N[1] = e1[2] * e2[0] - e1[0] * e2[2], N[2] = e1[0] * e2[1] - e1[1] * e2[0];
		normalize3( N );
		if (N[0] * D[0] + N[1] * D[1] + N[2] * D[2] > 0) N[0] = -N[0], N[1] = -N[1], N[2] = -N[2];
		if (hit && N[0] * R[0] + N[1] * R[1] + N[2] * R[2] < 0) R[0] = -R[0], R[1] = -R[1], R[2] = -R[2];
		float Oo[3];
		for (int a = 0; a < 3; a++) Oo[a] = I[a] + 0.001f * R[a];
		put_ray( out + (uint64_t)k * 128, Oo, R, BVH_FAR );
	}
}

/* reset hit.t / u / v / prim of every record (tiny_bvh_speedtest.cpp re-arms its batches the same way between passes) */
void tbvh_gen_reset_hits( char* rays, const uint64_t count, const float tmax )
{
	#pragma omp parallel for schedule( static )
	for (int64_t k = 0; k < (int64_t)count; k++)
	{
		float* f = (float*)(rays + (uint64_t)k * 128);
		f[12] = tmax, f[13] = 0, f[14] = 0, ((uint32_t*)f)[15] = 0;
	}
}

/* scatter a packed (t,u,v,prim) array into the records' hit fields */
void tbvh_gen_store_hits( char* rays, const float* hits, const uint64_t count )
{
	#pragma omp parallel for schedule( static )
	for (int64_t k = 0; k < (int64_t)count; k++) memcpy( rays + (uint64_t)k * 128 + 48, hits + (uint64_t)k * 4, 16 );
}
