// oracle/ref_api.h -- TEST INFRASTRUCTURE. C API of oracle/_ref/libref.so (unmodified reference
// + shim + cshare_spec) and of the call log shared with the restated oracle.
#pragma once
#include <stdint.h>
#include "../include/dmpp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
// one CShare::SearchObstacle evaluation ("one trajectory scored")
typedef struct ref_call {
    double lat_min, lat_max;
    double dis_lat, dis_lng;
    int32_t n_path;
    int16_t ob_index;
    uint16_t pathid;
    uint8_t found;
    uint8_t pad[7];
} ref_call;

int ref_set_map(const dp_map_desc* m);
long long ref_run_batch(int n, int cycles, int max_obs, const dp_scene_hdr* hdr, const double* ox,
                        const double* oy, dp_plan_record* rec, double* path_xy, double* path_ll,
                        ref_call* calls, int32_t* n_calls, int calls_cap, dp_carry* carry_out,
                        double* last_path_out, double* seconds, long long* traj_scored);
long long ref_search_calls(void);
int ref_sizeof(int which);
void share_set_datum(double lat0, double lng0, double k_lat, double k_lng);
#ifdef __cplusplus
}
#endif
