// Fake libobs vec3 (parity-oracle test infrastructure; see obs-module.h).
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
struct vec3 {
    union {
        struct { float x, y, z, w; };
        float ptr[4];
    };
};
static inline void vec3_set(struct vec3 *dst, float x, float y, float z) { dst->x = x; dst->y = y; dst->z = z; dst->w = 0.0f; }
static inline void vec3_copy(struct vec3 *dst, const struct vec3 *v) { *dst = *v; }
static inline void vec3_add(struct vec3 *dst, const struct vec3 *a, const struct vec3 *b)
{
    dst->x = a->x + b->x; dst->y = a->y + b->y; dst->z = a->z + b->z; dst->w = 0.0f;
}
#ifdef __cplusplus
}
#endif
