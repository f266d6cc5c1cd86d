// Fake libobs vec4 (parity-oracle test infrastructure; see obs-module.h).
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
struct vec4 {
    union {
        struct { float x, y, z, w; };
        float ptr[4];
    };
};
static inline void vec4_set(struct vec4 *dst, float x, float y, float z, float w) { dst->x = x; dst->y = y; dst->z = z; dst->w = w; }
#ifdef __cplusplus
}
#endif
