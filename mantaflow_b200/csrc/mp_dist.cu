// placeholder, replaced below
#include "mp_common.cuh"
extern "C" {
int mp_dist_unique_id(void*) { MP_FAIL(MP_ERR_UNSUPPORTED, "multi-GPU not built yet"); }
int mp_dist_init(mp_context*, int, int, const void*, const char*) { MP_FAIL(MP_ERR_UNSUPPORTED, "multi-GPU not built yet"); }
int mp_dist_shutdown(mp_context*) { return MP_OK; }
int mp_dist_slab(int sz, int rank, int world, int* k0, int* k1) { int base = sz / world, rem = sz % world; *k0 = rank * base + (rank < rem ? rank : rem); *k1 = *k0 + base + (rank < rem ? 1 : 0); return MP_OK; }
}
