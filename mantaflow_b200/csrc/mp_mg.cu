// placeholder, replaced below
#include "mp_common.cuh"
struct mp_mg { mp_context* ctx; };
int mp_mg_precond_init(mp_mg*, const mp_grid*, const mp_grid*, const mp_grid*, const mp_grid*, double) { MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg not built yet"); }
int mp_mg_precond_apply(mp_mg*, mp_grid*, const mp_grid*, const int*) { MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg not built yet"); }
extern "C" {
int mp_mg_create(mp_context*, int, int, int, int, mp_mg**) { MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg not built yet"); }
int mp_mg_destroy(mp_mg* mg) { delete mg; return MP_OK; }
int mp_mg_set_a(mp_mg*, const mp_grid*, const mp_grid*, const mp_grid*, const mp_grid*) { MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg not built yet"); }
int mp_mg_set_rhs(mp_mg*, const mp_grid*) { MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg not built yet"); }
int mp_mg_is_a_set(const mp_mg*, int* s) { *s = 0; return MP_OK; }
int mp_mg_do_vcycle(mp_mg*, mp_grid*, const mp_grid*, double*) { MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg not built yet"); }
int mp_mg_set_coarsest_level_accuracy(mp_mg*, double) { return MP_OK; }
int mp_mg_set_smoothing(mp_mg*, int, int) { return MP_OK; }
int mp_mg_num_levels(const mp_mg*, int* l) { *l = 0; return MP_OK; }
int mp_mg_level_info(const mp_mg*, int, int*, int*, int*, int*) { MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg not built yet"); }
int mp_mg_download(const mp_mg*, int, const char*, void*) { MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg not built yet"); }
}
