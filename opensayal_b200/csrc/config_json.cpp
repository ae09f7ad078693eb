// config_json.cpp — OpenSayal.conf.json reader for the simulation subset of the reference's Config.
//
// Mirrors ConfigParser::parse() (/root/reference/src/config_parser.cpp:15-185) and the lookup rule of
// ConfigParser::get_or (/root/reference/inc/config_parser.hpp:131-151):
//   1. if the current object contains the WHOLE key literally, take it;
//   2. else split at the FIRST dot, descend into the parent object and retry with the remainder;
//   3. anything that goes wrong below the top level silently yields the default.
// A wrong value type on a top-level literal key propagates in the reference (uncaught nlohmann
// type_error => abort); here it is SAYAL_EPARSE.  The reference depends on nlohmann/json (vendored
// submodule @568b708) only to parse numbers/bools; this ~200-line reader covers what the keys need.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "sayal.h"
#include "sayal_internal.h"

namespace {

struct JValue {
  enum Kind { Null, Bool, Int, Float, String, Array, Object } kind = Null;
  bool b = false;
  long long i = 0;
  double d = 0;
  std::string s;
  std::vector<std::pair<std::string, std::shared_ptr<JValue>>> members;  // object, in file order
  std::vector<std::shared_ptr<JValue>> items;
  const JValue* find(const std::string& key) const {
    if (kind != Object) return nullptr;
    const JValue* hit = nullptr;  // duplicate keys: the last one wins, as in nlohmann
    for (auto& m : members)
      if (m.first == key) hit = m.second.get();
    return hit;
  }
};

struct Parser {
  const char* p;
  const char* end;
  std::string err;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
  }
  bool fail(const char* m) {
    if (err.empty()) err = m;
    return false;
  }
  bool parse_string(std::string& out) {
    if (p >= end || *p != '"') return fail("expected string");
    p++;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        p++;
        if (p >= end) return fail("bad escape");
        switch (*p) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (end - p < 5) return fail("bad \\u escape");
            unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            p += 4;
            break;
          }
          default: out += *p;
        }
        p++;
      } else {
        out += *p++;
      }
    }
    if (p >= end) return fail("unterminated string");
    p++;
    return true;
  }
  bool parse_value(JValue& v, int depth) {
    if (depth > 64) return fail("nesting too deep");
    ws();
    if (p >= end) return fail("unexpected end of input");
    if (*p == '{') {
      v.kind = JValue::Object;
      p++;
      ws();
      if (p < end && *p == '}') { p++; return true; }
      while (true) {
        ws();
        std::string key;
        if (!parse_string(key)) return false;
        ws();
        if (p >= end || *p != ':') return fail("expected ':'");
        p++;
        auto child = std::make_shared<JValue>();
        if (!parse_value(*child, depth + 1)) return false;
        v.members.emplace_back(std::move(key), child);
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') { p++; return true; }
        return fail("expected ',' or '}'");
      }
    }
    if (*p == '[') {
      v.kind = JValue::Array;
      p++;
      ws();
      if (p < end && *p == ']') { p++; return true; }
      while (true) {
        auto child = std::make_shared<JValue>();
        if (!parse_value(*child, depth + 1)) return false;
        v.items.push_back(child);
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') { p++; return true; }
        return fail("expected ',' or ']'");
      }
    }
    if (*p == '"') {
      v.kind = JValue::String;
      return parse_string(v.s);
    }
    if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JValue::Bool; v.b = true; p += 4; return true; }
    if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JValue::Bool; v.b = false; p += 5; return true; }
    if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = JValue::Null; p += 4; return true; }
    // number
    const char* q = p;
    bool is_float = false;
    if (q < end && *q == '-') q++;
    if (q >= end || !isdigit((unsigned char)*q)) return fail("unexpected character");
    while (q < end && isdigit((unsigned char)*q)) q++;
    if (q < end && *q == '.') { is_float = true; q++; while (q < end && isdigit((unsigned char)*q)) q++; }
    if (q < end && (*q == 'e' || *q == 'E')) {
      is_float = true; q++;
      if (q < end && (*q == '+' || *q == '-')) q++;
      while (q < end && isdigit((unsigned char)*q)) q++;
    }
    std::string tok(p, q);
    if (is_float) { v.kind = JValue::Float; v.d = strtod(tok.c_str(), nullptr); }
    else { v.kind = JValue::Int; v.i = strtoll(tok.c_str(), nullptr, 10); v.d = (double)v.i; }
    p = q;
    return true;
  }
};

// nlohmann's implicit conversion rules for `T x = json_value;`
//   int / float targets accept integer, float and boolean values (arithmetic from_json);
//   bool targets accept only booleans.
enum ConvResult { CONV_OK, CONV_TYPE_ERROR };

ConvResult convert(const JValue& v, int* out) {
  if (v.kind == JValue::Int) { *out = (int)v.i; return CONV_OK; }
  if (v.kind == JValue::Float) { *out = (int)v.d; return CONV_OK; }
  if (v.kind == JValue::Bool) { *out = v.b ? 1 : 0; return CONV_OK; }
  return CONV_TYPE_ERROR;
}
ConvResult convert(const JValue& v, float* out) {
  if (v.kind == JValue::Int) { *out = (float)v.i; return CONV_OK; }
  if (v.kind == JValue::Float) { *out = (float)v.d; return CONV_OK; }
  if (v.kind == JValue::Bool) { *out = v.b ? 1.f : 0.f; return CONV_OK; }
  return CONV_TYPE_ERROR;
}
struct BoolT { int32_t v; };
ConvResult convert(const JValue& v, BoolT* out) {
  if (v.kind == JValue::Bool) { out->v = v.b ? 1 : 0; return CONV_OK; }
  return CONV_TYPE_ERROR;
}

// get_or (config_parser.hpp:131-151).  `top` marks the outermost frame, whose type errors are fatal.
template <typename T>
bool get_or(const JValue& obj, const std::string& key, T* out, bool top, bool* fatal) {
  if (const JValue* hit = obj.find(key)) {
    if (convert(*hit, out) == CONV_OK) return true;
    if (top) *fatal = true;
    return false;  // nested: the parent's catch(...) swallows it => default
  }
  if (key.empty()) return false;
  size_t dot = key.find('.');
  if (dot == std::string::npos) return false;
  const JValue* parent = obj.find(key.substr(0, dot));
  if (!parent) return false;  // .at(parent) throws, caught => default
  return get_or(*parent, key.substr(dot + 1), out, false, fatal);
}

}  // namespace

extern "C" int sayal_config_defaults(int32_t width, int32_t height, sayal_config* c) {
  if (!c) return sayal::set_error(SAYAL_EINVAL, "sayal_config_defaults: null config");
  std::memset(c, 0, sizeof(*c));
  c->width = width;
  c->height = height;
  c->cell_size = 1.0f;
  c->enable_drain = 1;
  c->enable_pressure = 0;
  c->enable_smoke = 1;
  c->enable_interactive = 0;
  c->proj_n = 50;
  c->proj_o = 1.9f;
  c->wt_pipe_height = height / 4;
  c->wt_pipe_length = 0;
  c->wt_smoke_length = 1;
  c->wt_smoke_height = height / 4;
  c->wt_smoke_count = 1;
  c->wt_speed = 0.0f;
  c->wt_smoke = 1.0f;
  c->g = 0.0f;
  c->d_t = 0.05f;
  c->enable_real_time = 0;
  c->real_time_multiplier = 1.0f;
  c->smoke_enable_decay = 0;
  c->smoke_decay_rate = 0.05f;
  c->obstacle_enable = 1;
  c->obstacle_center_x = width / 2;
  c->obstacle_center_y = height / 2;
  c->obstacle_radius = std::min(height, width) / 30.0f;
  c->density = 1.0f;
  c->drag_coeff = 0.0f;
  c->viscosity = 0.001f;
  c->block_size_x = 64;
  c->block_size_y = 1;
  return SAYAL_OK;
}

extern "C" int sayal_config_parse(const char* text, size_t len, sayal_config* c) {
  if (!text || !c) return sayal::set_error(SAYAL_EINVAL, "sayal_config_parse: null argument");
  Parser ps{text, text + len, {}};
  JValue root;
  if (!ps.parse_value(root, 0)) return sayal::set_error(SAYAL_EPARSE, ("config: " + ps.err).c_str());
  ps.ws();
  if (ps.p != ps.end) return sayal::set_error(SAYAL_EPARSE, "config: trailing characters after JSON value");

  bool fatal = false;
  int h = 1080, w = 1920;  // config_parser.cpp:21-22
  get_or(root, "sim.height", &h, true, &fatal);
  get_or(root, "sim.width", &w, true, &fatal);
  sayal_config_defaults(w, h, c);

#define GET_I(key, field) get_or(root, key, &c->field, true, &fatal)
#define GET_F(key, field) get_or(root, key, &c->field, true, &fatal)
#define GET_B(key, field)                                   \
  do {                                                      \
    BoolT t{c->field};                                      \
    if (get_or(root, key, &t, true, &fatal)) c->field = t.v; \
  } while (0)
  GET_I("thread.cuda.block_size_x", block_size_x);
  GET_I("thread.cuda.block_size_y", block_size_y);
  GET_F("sim.cell_size", cell_size);
  GET_B("sim.enable_drain", enable_drain);
  GET_B("sim.enable_pressure", enable_pressure);
  GET_B("sim.enable_smoke", enable_smoke);
  GET_B("sim.enable_interactive", enable_interactive);
  GET_I("sim.projection.n", proj_n);
  GET_F("sim.projection.o", proj_o);
  GET_I("sim.wind_tunnel.pipe_height", wt_pipe_height);
  GET_I("sim.wind_tunnel.smoke_length", wt_smoke_length);
  GET_I("sim.wind_tunnel.smoke_height", wt_smoke_height);
  GET_I("sim.wind_tunnel.smoke_count", wt_smoke_count);
  GET_F("sim.wind_tunnel.speed", wt_speed);
  GET_F("sim.wind_tunnel.smoke", wt_smoke);
  GET_F("sim.physics.g", g);
  GET_F("sim.time.d_t", d_t);
  GET_B("sim.time.enable_real_time", enable_real_time);
  GET_F("sim.time.real_time_multiplier", real_time_multiplier);
  GET_B("sim.smoke.enable_decay", smoke_enable_decay);
  GET_F("sim.smoke.decay_rate", smoke_decay_rate);
  GET_B("sim.obstacle.enable", obstacle_enable);
  GET_I("sim.obstacle.center_x", obstacle_center_x);
  GET_I("sim.obstacle.center_y", obstacle_center_y);
  GET_F("sim.obstacle.radius", obstacle_radius);
  GET_F("fluid.density", density);
  GET_F("fluid.drag_coeff", drag_coeff);
  GET_F("fluid.viscosity", viscosity);
#undef GET_I
#undef GET_F
#undef GET_B
  if (fatal) return sayal::set_error(SAYAL_EPARSE, "config: a top-level key holds a value of the wrong type");
  return SAYAL_OK;
}

extern "C" int sayal_config_load(const char* path, sayal_config* c) {
  if (!path || !c) return sayal::set_error(SAYAL_EINVAL, "sayal_config_load: null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return sayal::set_error(SAYAL_EIO, (std::string("cannot open config file ") + path).c_str());
  std::string text;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
  fclose(f);
  return sayal_config_parse(text.data(), text.size(), c);
}

// ---- the renderer's subset of Config (config_parser.cpp:40, 120-181) ------------------------------------------
extern "C" int sayal_visual_defaults(sayal_visual* v) {
  if (!v) return sayal::set_error(SAYAL_EINVAL, "sayal_visual_defaults: null argument");
  std::memset(v, 0, sizeof(*v));
  v->cell_pixel_size = 1;
  v->arrows_enable = 0;
  v->arrows_distance = 20;
  v->arrows_length_multiplier = 0.1f;
  v->arrows_disable_threshold = 0.0f;
  v->arrows_head_length = 5;
  v->path_line_enable = 0;
  v->path_line_length = 20;
  v->path_line_distance = 20;
  v->arrows_color[3] = 255;
  v->path_line_color[3] = 255;
  return SAYAL_OK;
}

extern "C" int sayal_visual_parse(const char* text, size_t len, sayal_visual* v) {
  if (!text || !v) return sayal::set_error(SAYAL_EINVAL, "sayal_visual_parse: null argument");
  Parser ps{text, text + len, {}};
  JValue root;
  if (!ps.parse_value(root, 0)) return sayal::set_error(SAYAL_EPARSE, ("config: " + ps.err).c_str());
  ps.ws();
  if (ps.p != ps.end) return sayal::set_error(SAYAL_EPARSE, "config: trailing characters after JSON value");
  sayal_visual_defaults(v);
  bool fatal = false;
#define VGET(key, field) get_or(root, key, &v->field, true, &fatal)
#define VGET_B(key, field)                                  \
  do {                                                      \
    BoolT t{v->field};                                      \
    if (get_or(root, key, &t, true, &fatal)) v->field = t.v; \
  } while (0)
  VGET("sim.cell_pixel_size", cell_pixel_size);
  VGET_B("visual.arrows.enable", arrows_enable);
  VGET("visual.arrows.distance", arrows_distance);
  VGET("visual.arrows.length_multiplier", arrows_length_multiplier);
  VGET("visual.arrows.disable_threshold", arrows_disable_threshold);
  VGET("visual.arrows.head_length", arrows_head_length);
  VGET_B("visual.path_line.enable", path_line_enable);
  VGET("visual.path_line.length", path_line_length);
  VGET("visual.path_line.distance", path_line_distance);
  static const char* rgba[4] = {"r", "g", "b", "a"};
  for (int k = 0; k < 4; k++) {
    get_or(root, std::string("visual.arrows.color.") + rgba[k], &v->arrows_color[k], true, &fatal);
    get_or(root, std::string("visual.path_line.color.") + rgba[k], &v->path_line_color[k], true, &fatal);
  }
#undef VGET
#undef VGET_B
  if (fatal) return sayal::set_error(SAYAL_EPARSE, "config: a top-level key holds a value of the wrong type");
  return SAYAL_OK;
}

extern "C" int sayal_visual_load(const char* path, sayal_visual* v) {
  if (!path || !v) return sayal::set_error(SAYAL_EINVAL, "sayal_visual_load: null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return sayal::set_error(SAYAL_EIO, (std::string("cannot open config file ") + path).c_str());
  std::string text;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
  fclose(f);
  return sayal_visual_parse(text.data(), text.size(), v);
}
