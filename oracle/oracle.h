/* oracle/oracle.h — TEST INFRASTRUCTURE.  C ABI of the CPU oracle: the same entry points as
 * include/mox.h with the prefix orc_ instead of mox_, so tests can feed one scene to both.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libmox.so) never does. */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include "../include/mox.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_ctx orc_ctx;

int orc_create(orc_ctx** out, int device_id /* ignored */);
void orc_destroy(orc_ctx*);
const char* orc_last_error(const orc_ctx*);
int orc_abi_version(void);
int orc_set_globals(orc_ctx*, uint32_t width, uint32_t height, uint32_t rayMaxDepth, float rayEpsilonT,
                    float rayMinIntensity, const float absorbColor[3], const float badColor[3],
                    const float bgColor[3]);
int orc_set_camera(orc_ctx*, const CamParams*);
int orc_set_rng_mode(orc_ctx*, int mode);
int orc_set_partition(orc_ctx*, uint32_t rank, uint32_t world, uint32_t tile);
int orc_add_texture_rgba32f(orc_ctx*, const float* texels, int w, int h, int* out_id);
int orc_add_sphere(orc_ctx*, const SphereParams*, int kind, const void* params, uint32_t* out_prim_id);
int orc_add_quad(orc_ctx*, const QuadParams*, int kind, const void* params, uint32_t* out_prim_id);
int orc_add_mesh(orc_ctx*, const float* vertices, size_t n_vertices, const float* normals, size_t n_normals,
                 const float* texcoords, size_t n_texcoords, const int32_t* vIdx, const int32_t* nIdx,
                 const int32_t* tIdx, size_t n_faces, int kind, const void* params, uint32_t* out_first_prim_id);
int orc_set_lights(orc_ctx*, const LightParams*, size_t n);
int orc_clear_scene(orc_ctx*);
int orc_build_accel(orc_ctx*, uint32_t flags, float* out_build_ms);
int orc_launch(orc_ctx*, int32_t randSeed);
int orc_render(orc_ctx*, uint32_t spp, uint32_t seed);
int orc_read_accum(orc_ctx*, float* dst_rgb);
int orc_map_accum(orc_ctx*, const float** out);
int orc_unmap_accum(orc_ctx*);
int orc_device_count(const orc_ctx*);
int orc_get_device_stats(orc_ctx*, int index, mox_stats* out);
int orc_read_accum_begin(orc_ctx*);
int orc_read_accum_end(orc_ctx*, const float** out);
int orc_clear_accum(orc_ctx*);
int orc_set_accum(orc_ctx*, const float* src_rgb, uint64_t launches);
int orc_update_sphere(orc_ctx*, uint32_t prim_id, const SphereParams*);
int orc_owned_pixels(orc_ctx*, uint32_t rank, uint64_t* out_n);
int orc_pack_owned(orc_ctx*, void* dst);            /* host pointers in the oracle */
int orc_unpack_owned(orc_ctx*, uint32_t rank, const void* src);
int orc_get_stats(orc_ctx*, mox_stats*);
int orc_trace_closest(orc_ctx*, const float* rays, size_t n, void* hits);
int orc_trace_shadow(orc_ctx*, const float* rays, size_t n, float* out_rgb);

/* oracle-only knobs */
int orc_set_threads(orc_ctx*, int n_threads);       /* 0 = hardware_concurrency */
int orc_set_brute_force(orc_ctx*, int on);          /* 1 = test every primitive (id ground truth) */
/* Which of the two draws of `light.u * rand(s) + light.v * rand(s)` (Material.cu:180) scales u:
 * 0 = the first (pinned convention, nvcc's order), 1 = the second (g++'s order; used only when
 * comparing against the reference compiled with g++ in oracle/_ref). */
int orc_set_quad_light_draw_order(orc_ctx*, int order);
/* unit-test hooks for the restated device functions */
uint32_t orc_tea16(uint32_t v0, uint32_t v1);
uint32_t orc_lcg(int32_t* seed);
float orc_rand(int32_t* seed);
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_disney_eval(const DisneyParams*, const float baseColor[3], const float N[3], const float L[3],
                     const float V[3], const float H[3], float out[3]);
float orc_disney_pdf(const DisneyParams*, const float N[3], const float L[3], const float V[3], const float H[3]);
void orc_disney_sample(int32_t* seed, const DisneyParams*, const float N[3], const float V[3], float L[3],
                       float H[3]);
void orc_refine_hitpoint(const float hit[3], const float dir[3], const float n[3], const float p[3],
                         float back[3], float front[3]);
int orc_refract(const float i[3], const float n[3], float ior, float out[3]);
float orc_fresnel(float cosI, float cosT, float ior);
void orc_rand_in_unit_sphere(int32_t* seed, float out[3]);
void orc_rand_in_unit_disk(int32_t* seed, float out[3]);
int32_t orc_fork_seed(int32_t parentSeed, int32_t parentDepth);
void orc_offset(const float hit[3], const float n[3], float out[3]);
float orc_gtr1(float NdotH, float a);
float orc_gtr2(float NdotH, float a);
float orc_gtr2_aniso(float NdotH, float HdotX, float HdotY, float ax, float ay);
float orc_schlick_fresnel(float u);
float orc_smith_ggx(float NdotV, float alphaG);
float orc_smith_ggx_aniso(float NdotV, float VdotX, float VdotY, float ax, float ay);
float orc_power_heuristic(float a, float b);
void orc_srgb2lin(const float v[3], float out[3]);
/* closest hit + the five intersection attributes (n x 15 floats); bbox-program bounds */
int orc_trace_closest_attrs(orc_ctx*, const float* rays, size_t n, void* hits, float* attrs);
int orc_prim_bounds(orc_ctx*, uint32_t prim, float out[6], int* valid);
#ifdef __cplusplus
}
#endif
#endif
