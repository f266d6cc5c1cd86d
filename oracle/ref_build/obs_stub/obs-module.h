// Fake libobs for the parity oracle (TEST INFRASTRUCTURE ONLY — not shipped, not measured).
//
// The reference plugin (/root/reference/src/*.cpp) is compiled UNMODIFIED against this header so
// that its own update()/capture_audio()/tick()/render() code builds every table and runs the
// spectrum path.  Only the slice of the libobs API that those translation units mention is
// declared; behaviour is the minimum needed to drive the hot path:
//   * obs_data_t            = string -> variant map (settings), filled by ref_harness.cpp
//   * os_gettime_ns()       = controllable fake clock (wfref_clock_ns)
//   * obs_get_audio_info()  = sample rate / speaker layout chosen by the harness
//   * audio_output_connect  = "succeeds" so the output-bus capture path is armed
//                             (src/source.cpp:685-703) and capture_audio() accepts data
//   * gs_* / properties     = inert
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdarg>

#ifdef __cplusplus
extern "C" {
#endif

// ---- logging / memory -------------------------------------------------------------------------
enum { LOG_ERROR = 100, LOG_WARNING = 200, LOG_INFO = 300, LOG_DEBUG = 400 };
void blog(int level, const char *fmt, ...);
void *bmalloc(size_t sz);
void *bzalloc(size_t sz);
void bfree(void *p);

// ---- audio ------------------------------------------------------------------------------------
#define MAX_AUDIO_CHANNELS 8
#define MAX_AV_PLANES 8
#define AUDIO_OUTPUT_FRAMES 1024

enum speaker_layout {
    SPEAKERS_UNKNOWN = 0,
    SPEAKERS_MONO,
    SPEAKERS_STEREO,
    SPEAKERS_2POINT1,
    SPEAKERS_4POINT0,
    SPEAKERS_4POINT1,
    SPEAKERS_5POINT1,
    SPEAKERS_7POINT1 = 8,
};

enum audio_format {
    AUDIO_FORMAT_UNKNOWN = 0,
    AUDIO_FORMAT_U8BIT,
    AUDIO_FORMAT_16BIT,
    AUDIO_FORMAT_32BIT,
    AUDIO_FORMAT_FLOAT,
    AUDIO_FORMAT_U8BIT_PLANAR,
    AUDIO_FORMAT_16BIT_PLANAR,
    AUDIO_FORMAT_32BIT_PLANAR,
    AUDIO_FORMAT_FLOAT_PLANAR,
};

struct audio_data {
    uint8_t *data[MAX_AV_PLANES];
    uint32_t frames;
    uint64_t timestamp;
};

struct obs_audio_info {
    uint32_t samples_per_sec;
    enum speaker_layout speakers;
};

struct audio_output_info {
    const char *name;
    uint32_t samples_per_sec;
    enum audio_format format;
    enum speaker_layout speakers;
};

struct audio_convert_info {
    uint32_t samples_per_sec;
    enum audio_format format;
    enum speaker_layout speakers;
    bool allow_clipping;
};

typedef struct audio_output audio_t;
typedef void (*audio_output_callback_t)(void *param, size_t mix_idx, struct audio_data *data);

static inline uint32_t get_audio_channels(enum speaker_layout speakers)
{
    switch(speakers) {
    case SPEAKERS_MONO: return 1;
    case SPEAKERS_STEREO: return 2;
    case SPEAKERS_2POINT1: return 3;
    case SPEAKERS_4POINT0: return 4;
    case SPEAKERS_4POINT1: return 5;
    case SPEAKERS_5POINT1: return 6;
    case SPEAKERS_7POINT1: return 8;
    default: return 0;
    }
}

// same integer arithmetic libobs documents: frames <-> ns at a sample rate
static inline uint64_t ns_to_audio_frames(size_t sample_rate, uint64_t ns)
{
    return (uint64_t)(((__uint128_t)ns * (__uint128_t)sample_rate) / 1000000000ull);
}
static inline uint64_t audio_frames_to_ns(size_t sample_rate, uint64_t frames)
{
    return (uint64_t)(((__uint128_t)frames * 1000000000ull) / (__uint128_t)sample_rate);
}

bool obs_get_audio_info(struct obs_audio_info *info);
audio_t *obs_get_audio(void);
const struct audio_output_info *audio_output_get_info(const audio_t *audio);
bool audio_output_connect(audio_t *audio, size_t mix_idx, const struct audio_convert_info *conversion,
                          audio_output_callback_t callback, void *param);
void audio_output_disconnect(audio_t *audio, size_t mix_idx, audio_output_callback_t callback, void *param);

// ---- video ------------------------------------------------------------------------------------
struct obs_video_info {
    uint32_t fps_num;
    uint32_t fps_den;
};
bool obs_get_video_info(struct obs_video_info *ovi);

// ---- sources ----------------------------------------------------------------------------------
typedef struct obs_source obs_source_t;
typedef struct obs_weak_source obs_weak_source_t;
typedef struct obs_data obs_data_t;
typedef struct obs_properties obs_properties_t;
typedef struct obs_property obs_property_t;
typedef struct gs_effect gs_effect_t;
typedef struct gs_technique gs_technique_t;
typedef struct gs_effect_param gs_eparam_t;
typedef struct gs_vertex_buffer gs_vertbuffer_t;
typedef struct gs_index_buffer gs_indexbuffer_t;

#define OBS_SOURCE_VIDEO (1 << 0)
#define OBS_SOURCE_AUDIO (1 << 1)
#define OBS_SOURCE_CUSTOM_DRAW (1 << 3)

enum obs_source_type { OBS_SOURCE_TYPE_INPUT = 0 };
enum obs_icon_type { OBS_ICON_TYPE_UNKNOWN = 0, OBS_ICON_TYPE_AUDIO_OUTPUT = 4 };

typedef void (*obs_source_audio_capture_t)(void *param, obs_source_t *source, const struct audio_data *audio_data,
                                           bool muted);

struct obs_source_info {
    const char *id;
    enum obs_source_type type;
    uint32_t output_flags;
    const char *(*get_name)(void *type_data);
    void *(*create)(obs_data_t *settings, obs_source_t *source);
    void (*destroy)(void *data);
    uint32_t (*get_width)(void *data);
    uint32_t (*get_height)(void *data);
    void (*get_defaults)(obs_data_t *settings);
    obs_properties_t *(*get_properties)(void *data);
    void (*update)(void *data, obs_data_t *settings);
    void (*show)(void *data);
    void (*hide)(void *data);
    void (*video_tick)(void *data, float seconds);
    void (*video_render)(void *data, gs_effect_t *effect);
    enum obs_icon_type icon_type;
};
void obs_register_source(const struct obs_source_info *info);

bool obs_source_showing(const obs_source_t *source);
uint32_t obs_source_get_output_flags(const obs_source_t *source);
const char *obs_source_get_name(const obs_source_t *source);
obs_source_t *obs_get_source_by_name(const char *name);
obs_weak_source_t *obs_source_get_weak_source(obs_source_t *source);
obs_source_t *obs_weak_source_get_source(obs_weak_source_t *weak);
void obs_weak_source_release(obs_weak_source_t *weak);
void obs_source_release(obs_source_t *source);
void obs_source_add_audio_capture_callback(obs_source_t *source, obs_source_audio_capture_t callback, void *param);
void obs_source_remove_audio_capture_callback(obs_source_t *source, obs_source_audio_capture_t callback, void *param);
void obs_enum_sources(bool (*enum_proc)(void *, obs_source_t *), void *param);

// ---- settings (obs_data) ----------------------------------------------------------------------
const char *obs_data_get_string(obs_data_t *data, const char *name);
long long obs_data_get_int(obs_data_t *data, const char *name);
double obs_data_get_double(obs_data_t *data, const char *name);
bool obs_data_get_bool(obs_data_t *data, const char *name);
void obs_data_set_default_string(obs_data_t *data, const char *name, const char *val);
void obs_data_set_default_int(obs_data_t *data, const char *name, long long val);
void obs_data_set_default_double(obs_data_t *data, const char *name, double val);
void obs_data_set_default_bool(obs_data_t *data, const char *name, bool val);

// ---- properties UI (inert) --------------------------------------------------------------------
enum obs_combo_type { OBS_COMBO_TYPE_INVALID, OBS_COMBO_TYPE_EDITABLE, OBS_COMBO_TYPE_LIST };
enum obs_combo_format { OBS_COMBO_FORMAT_INVALID, OBS_COMBO_FORMAT_INT, OBS_COMBO_FORMAT_FLOAT, OBS_COMBO_FORMAT_STRING };
typedef bool (*obs_property_modified_t)(obs_properties_t *props, obs_property_t *property, obs_data_t *settings);

obs_properties_t *obs_properties_create(void);
obs_property_t *obs_properties_get(obs_properties_t *props, const char *property);
obs_property_t *obs_properties_add_bool(obs_properties_t *props, const char *name, const char *description);
obs_property_t *obs_properties_add_int(obs_properties_t *props, const char *name, const char *description, int min,
                                       int max, int step);
obs_property_t *obs_properties_add_int_slider(obs_properties_t *props, const char *name, const char *description,
                                              int min, int max, int step);
obs_property_t *obs_properties_add_float_slider(obs_properties_t *props, const char *name, const char *description,
                                                double min, double max, double step);
obs_property_t *obs_properties_add_list(obs_properties_t *props, const char *name, const char *description,
                                        enum obs_combo_type type, enum obs_combo_format format);
obs_property_t *obs_properties_add_color(obs_properties_t *props, const char *name, const char *description);
size_t obs_property_list_add_string(obs_property_t *p, const char *name, const char *val);
void obs_property_list_item_disable(obs_property_t *p, size_t idx, bool disabled);
void obs_property_set_modified_callback(obs_property_t *p, obs_property_modified_t modified);
void obs_property_set_long_description(obs_property_t *p, const char *long_description);
void obs_property_set_visible(obs_property_t *p, bool visible);
void obs_property_set_enabled(obs_property_t *p, bool enabled);
bool obs_property_visible(obs_property_t *p);
void obs_property_int_set_limits(obs_property_t *p, int min, int max, int step);
void obs_property_int_set_suffix(obs_property_t *p, const char *suffix);
void obs_property_float_set_suffix(obs_property_t *p, const char *suffix);

// ---- graphics (inert) -------------------------------------------------------------------------
struct vec2 { float x, y; };
static inline void vec2_set(struct vec2 *dst, float x, float y) { dst->x = x; dst->y = y; }

#include "graphics/vec3.h"
#include "graphics/vec4.h"

struct gs_tvertarray {
    size_t width;
    void *array;
};
struct gs_vb_data {
    size_t num;
    struct vec3 *points;
    struct vec3 *normals;
    struct vec3 *tangents;
    uint32_t *colors;
    size_t num_tex;
    struct gs_tvertarray *tvarray;
};
enum gs_draw_mode { GS_POINTS, GS_LINES, GS_LINESTRIP, GS_TRIS, GS_TRISTRIP };
#define GS_DYNAMIC (1 << 1)

void obs_enter_graphics(void);
void obs_leave_graphics(void);
struct gs_vb_data *gs_vbdata_create(void);
gs_vertbuffer_t *gs_vertexbuffer_create(struct gs_vb_data *data, uint32_t flags);
void gs_vertexbuffer_destroy(gs_vertbuffer_t *vertbuffer);
void gs_vertexbuffer_flush(gs_vertbuffer_t *vertbuffer);
struct gs_vb_data *gs_vertexbuffer_get_data(const gs_vertbuffer_t *vertbuffer);
void gs_load_vertexbuffer(gs_vertbuffer_t *vertbuffer);
void gs_load_indexbuffer(gs_indexbuffer_t *indexbuffer);
void gs_draw(enum gs_draw_mode draw_mode, uint32_t start_vert, uint32_t num_verts);
gs_effect_t *gs_effect_create_from_file(const char *file, char **error_string);
void gs_effect_destroy(gs_effect_t *effect);
gs_technique_t *gs_effect_get_technique(const gs_effect_t *effect, const char *name);
gs_eparam_t *gs_effect_get_param_by_name(const gs_effect_t *effect, const char *name);
size_t gs_technique_begin(gs_technique_t *technique);
void gs_technique_end(gs_technique_t *technique);
bool gs_technique_begin_pass(gs_technique_t *technique, size_t pass);
void gs_technique_end_pass(gs_technique_t *technique);
void gs_effect_set_bool(gs_eparam_t *param, bool val);
void gs_effect_set_float(gs_eparam_t *param, float val);
void gs_effect_set_vec2(gs_eparam_t *param, const struct vec2 *val);
void gs_effect_set_vec4(gs_eparam_t *param, const struct vec4 *val);

// ---- module -----------------------------------------------------------------------------------
const char *obs_module_text(const char *lookup_string);
char *obs_module_file(const char *file);
#define OBS_DECLARE_MODULE()
#define OBS_MODULE_USE_DEFAULT_LOCALE(name, locale)
#define MODULE_EXPORT

#ifdef __cplusplus
}
#endif
