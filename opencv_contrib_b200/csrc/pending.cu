// pending.cu -- entry points declared in include/b200flow.h whose engines have not landed yet.
// They fail loudly (no fallback).  Each block is deleted when the engine's own .cu arrives.
#include "common.cuh"

extern "C" {

void b2f_brox_default_params(b2f_brox_params *p) {
    if (!p) return;
    p->alpha = 0.197; p->gamma = 50.0; p->scale_factor = 0.8;
    p->inner_iterations = 5; p->outer_iterations = 150; p->solver_iterations = 10;
}
int b2f_brox_create(const b2f_brox_params *, b2f_handle **out) {
    if (out) *out = nullptr;
    return B2F_UNSUPPORTED_TYPE;
}

void b2f_denselk_default_params(b2f_denselk_params *p) {
    if (!p) return;
    p->win_width = 13; p->win_height = 13; p->max_level = 3; p->iters = 30; p->use_initial_flow = 0;
}
int b2f_denselk_create(const b2f_denselk_params *, b2f_handle **out) {
    if (out) *out = nullptr;
    return B2F_UNSUPPORTED_TYPE;
}

}  // extern "C"
