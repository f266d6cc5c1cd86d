// Fake libobs implementation for the parity oracle (TEST INFRASTRUCTURE ONLY).
// See obs_stub/obs-module.h for what is modelled and why.  Everything graphics/UI is inert; the
// pieces the spectrum path depends on (settings map, fake clock, audio info, output-bus connect)
// are controllable from ref_harness.cpp through the wfstub_* hooks declared in obs_stub_hooks.h.
#include <obs-module.h>
#include <util/platform.h>
#include "obs_stub_hooks.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <variant>
#include <vector>

// ---------------------------------------------------------------------------------------------
// controllable state (thread_local so that several reference instances can be driven from
// several host threads for the CPU baseline without sharing a clock)
// ---------------------------------------------------------------------------------------------
static thread_local uint64_t t_clock_ns = 10ull * 1000000000ull;
static thread_local uint32_t t_sample_rate = 48000;
static thread_local speaker_layout t_speakers = SPEAKERS_STEREO;
static thread_local uint32_t t_fps_num = 60, t_fps_den = 1;
static thread_local bool t_showing = true;
static int g_log_level = LOG_ERROR;   // print only errors by default
static obs_source_info g_registered{};
static bool g_have_registered = false;

extern "C" {

void wfstub_set_clock_ns(uint64_t ns) { t_clock_ns = ns; }
uint64_t wfstub_get_clock_ns(void) { return t_clock_ns; }
void wfstub_set_audio(uint32_t sample_rate, int channels)
{
    t_sample_rate = sample_rate;
    t_speakers = (channels >= 2) ? SPEAKERS_STEREO : (channels == 1 ? SPEAKERS_MONO : SPEAKERS_UNKNOWN);
}
void wfstub_set_fps(uint32_t num, uint32_t den) { t_fps_num = num; t_fps_den = den; }
void wfstub_set_showing(bool s) { t_showing = s; }
void wfstub_set_log_level(int lvl) { g_log_level = lvl; }
const obs_source_info *wfstub_registered_info(void) { return g_have_registered ? &g_registered : nullptr; }

uint64_t os_gettime_ns(void) { return t_clock_ns; }

// ---- logging / memory ----
void blog(int level, const char *fmt, ...)
{
    if(level > g_log_level)
        return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
}
void *bmalloc(size_t sz) { return malloc(sz ? sz : 1); }
void *bzalloc(size_t sz) { return calloc(1, sz ? sz : 1); }
void bfree(void *p) { free(p); }

// ---- audio / video info ----
struct audio_output { audio_output_info info; };
static thread_local audio_output t_audio{};

bool obs_get_audio_info(struct obs_audio_info *info)
{
    info->samples_per_sec = t_sample_rate;
    info->speakers = t_speakers;
    return true;
}
audio_t *obs_get_audio(void)
{
    t_audio.info.name = "wfstub";
    t_audio.info.samples_per_sec = t_sample_rate;
    t_audio.info.format = AUDIO_FORMAT_FLOAT_PLANAR;
    t_audio.info.speakers = t_speakers;
    return &t_audio;
}
const struct audio_output_info *audio_output_get_info(const audio_t *audio) { return &audio->info; }
bool audio_output_connect(audio_t *, size_t, const struct audio_convert_info *, audio_output_callback_t, void *) { return true; }
void audio_output_disconnect(audio_t *, size_t, audio_output_callback_t, void *) {}
bool obs_get_video_info(struct obs_video_info *ovi)
{
    ovi->fps_num = t_fps_num;
    ovi->fps_den = t_fps_den;
    return true;
}

// ---- sources ----
void obs_register_source(const struct obs_source_info *info)
{
    g_registered = *info;
    g_have_registered = true;
}
bool obs_source_showing(const obs_source_t *) { return t_showing; }
uint32_t obs_source_get_output_flags(const obs_source_t *) { return 0; }
const char *obs_source_get_name(const obs_source_t *) { return "wfstub"; }
obs_source_t *obs_get_source_by_name(const char *) { return nullptr; }
obs_weak_source_t *obs_source_get_weak_source(obs_source_t *) { return nullptr; }
obs_source_t *obs_weak_source_get_source(obs_weak_source_t *) { return nullptr; }
void obs_weak_source_release(obs_weak_source_t *) {}
void obs_source_release(obs_source_t *) {}
void obs_source_add_audio_capture_callback(obs_source_t *, obs_source_audio_capture_t, void *) {}
void obs_source_remove_audio_capture_callback(obs_source_t *, obs_source_audio_capture_t, void *) {}
void obs_enum_sources(bool (*)(void *, obs_source_t *), void *) {}

} // extern "C"

// ---- settings ----
struct obs_data {
    using Val = std::variant<long long, double, bool, std::string>;
    std::map<std::string, Val> vals, defaults;
    const Val *find(const char *name) const
    {
        auto it = vals.find(name);
        if(it != vals.end())
            return &it->second;
        it = defaults.find(name);
        if(it != defaults.end())
            return &it->second;
        return nullptr;
    }
};

static double as_double(const obs_data::Val &v)
{
    if(auto p = std::get_if<double>(&v)) return *p;
    if(auto p = std::get_if<long long>(&v)) return (double)*p;
    if(auto p = std::get_if<bool>(&v)) return *p ? 1.0 : 0.0;
    return 0.0;
}
static long long as_int(const obs_data::Val &v)
{
    if(auto p = std::get_if<long long>(&v)) return *p;
    if(auto p = std::get_if<double>(&v)) return (long long)*p;
    if(auto p = std::get_if<bool>(&v)) return *p ? 1 : 0;
    return 0;
}

extern "C" {

obs_data_t *wfstub_data_create(void) { return new obs_data(); }
void wfstub_data_destroy(obs_data_t *d) { delete d; }
void wfstub_data_set_int(obs_data_t *d, const char *k, long long v) { d->vals[k] = v; }
void wfstub_data_set_double(obs_data_t *d, const char *k, double v) { d->vals[k] = v; }
void wfstub_data_set_bool(obs_data_t *d, const char *k, bool v) { d->vals[k] = v; }
void wfstub_data_set_string(obs_data_t *d, const char *k, const char *v) { d->vals[k] = std::string(v); }

const char *obs_data_get_string(obs_data_t *data, const char *name)
{
    auto v = data->find(name);
    if(v)
        if(auto p = std::get_if<std::string>(v))
            return p->c_str();
    return "";
}
long long obs_data_get_int(obs_data_t *data, const char *name)
{
    auto v = data->find(name);
    return v ? as_int(*v) : 0;
}
double obs_data_get_double(obs_data_t *data, const char *name)
{
    auto v = data->find(name);
    return v ? as_double(*v) : 0.0;
}
bool obs_data_get_bool(obs_data_t *data, const char *name)
{
    auto v = data->find(name);
    return v ? (as_int(*v) != 0) : false;
}
void obs_data_set_default_string(obs_data_t *data, const char *name, const char *val) { data->defaults[name] = std::string(val); }
void obs_data_set_default_int(obs_data_t *data, const char *name, long long val) { data->defaults[name] = val; }
void obs_data_set_default_double(obs_data_t *data, const char *name, double val) { data->defaults[name] = val; }
void obs_data_set_default_bool(obs_data_t *data, const char *name, bool val) { data->defaults[name] = val; }

// ---- properties UI: inert, but hand back stable non-null handles ----
} // extern "C"
struct obs_property { bool visible = true; };
struct obs_properties { std::map<std::string, obs_property> props; };
extern "C" {
obs_properties_t *obs_properties_create(void) { return new obs_properties(); }
obs_property_t *obs_properties_get(obs_properties_t *props, const char *property) { return &props->props[property]; }
obs_property_t *obs_properties_add_bool(obs_properties_t *p, const char *n, const char *) { return &p->props[n]; }
obs_property_t *obs_properties_add_int(obs_properties_t *p, const char *n, const char *, int, int, int) { return &p->props[n]; }
obs_property_t *obs_properties_add_int_slider(obs_properties_t *p, const char *n, const char *, int, int, int) { return &p->props[n]; }
obs_property_t *obs_properties_add_float_slider(obs_properties_t *p, const char *n, const char *, double, double, double) { return &p->props[n]; }
obs_property_t *obs_properties_add_list(obs_properties_t *p, const char *n, const char *, enum obs_combo_type, enum obs_combo_format) { return &p->props[n]; }
obs_property_t *obs_properties_add_color(obs_properties_t *p, const char *n, const char *) { return &p->props[n]; }
size_t obs_property_list_add_string(obs_property_t *, const char *, const char *) { return 0; }
void obs_property_list_item_disable(obs_property_t *, size_t, bool) {}
void obs_property_set_modified_callback(obs_property_t *, obs_property_modified_t) {}
void obs_property_set_long_description(obs_property_t *, const char *) {}
void obs_property_set_visible(obs_property_t *p, bool visible) { p->visible = visible; }
void obs_property_set_enabled(obs_property_t *, bool) {}
bool obs_property_visible(obs_property_t *p) { return p->visible; }
void obs_property_int_set_limits(obs_property_t *, int, int, int) {}
void obs_property_int_set_suffix(obs_property_t *, const char *) {}
void obs_property_float_set_suffix(obs_property_t *, const char *) {}

// ---- graphics: inert, vertex buffer keeps its CPU-side data so render() can fill it ----
} // extern "C"
struct gs_vertex_buffer { gs_vb_data *data; };
struct gs_effect { int dummy; };
struct gs_technique { int dummy; };
struct gs_effect_param { int dummy; };
static gs_effect g_effect;
static gs_technique g_tech;
static gs_effect_param g_param;
extern "C" {
void obs_enter_graphics(void) {}
void obs_leave_graphics(void) {}
struct gs_vb_data *gs_vbdata_create(void) { return (gs_vb_data *)calloc(1, sizeof(gs_vb_data)); }
gs_vertbuffer_t *gs_vertexbuffer_create(struct gs_vb_data *data, uint32_t) { return new gs_vertex_buffer{data}; }
void gs_vertexbuffer_destroy(gs_vertbuffer_t *vb)
{
    if(!vb)
        return;
    if(vb->data)
    {
        free(vb->data->points);
        if(vb->data->tvarray)
        {
            free(vb->data->tvarray->array);
            free(vb->data->tvarray);
        }
        free(vb->data);
    }
    delete vb;
}
void gs_vertexbuffer_flush(gs_vertbuffer_t *) {}
struct gs_vb_data *gs_vertexbuffer_get_data(const gs_vertbuffer_t *vb) { return vb->data; }
void gs_load_vertexbuffer(gs_vertbuffer_t *) {}
void gs_load_indexbuffer(gs_indexbuffer_t *) {}
void gs_draw(enum gs_draw_mode, uint32_t, uint32_t) {}
gs_effect_t *gs_effect_create_from_file(const char *, char **) { return &g_effect; }
void gs_effect_destroy(gs_effect_t *) {}
gs_technique_t *gs_effect_get_technique(const gs_effect_t *, const char *) { return &g_tech; }
gs_eparam_t *gs_effect_get_param_by_name(const gs_effect_t *, const char *) { return &g_param; }
size_t gs_technique_begin(gs_technique_t *) { return 1; }
void gs_technique_end(gs_technique_t *) {}
bool gs_technique_begin_pass(gs_technique_t *, size_t) { return true; }
void gs_technique_end_pass(gs_technique_t *) {}
void gs_effect_set_bool(gs_eparam_t *, bool) {}
void gs_effect_set_float(gs_eparam_t *, float) {}
void gs_effect_set_vec2(gs_eparam_t *, const struct vec2 *) {}
void gs_effect_set_vec4(gs_eparam_t *, const struct vec4 *) {}

// ---- module ----
const char *obs_module_text(const char *lookup_string) { return lookup_string; }
char *obs_module_file(const char *file)
{
    size_t n = strlen(file) + 1;
    char *p = (char *)bmalloc(n);
    memcpy(p, file, n);
    return p;
}

} // extern "C"
