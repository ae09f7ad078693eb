// sayal_run — headless driver over the C ABI; the replacement for the SDL-bound loop of
// /root/reference/src/main.cu:42-107 (ConfigParser().parse(); Fluid fluid(config); loop { fluid.update }).
//
//   sayal_run [--config OpenSayal.conf.json] [--steps N] [--device D] [--dump PREFIX] [--frames PREFIX]
//             [--every K] [--plain] [--temporal-block T] [--no-graph] [--real-time]
//
// Like the reference it reads ./OpenSayal.conf.json by default (config_parser.cpp:13) and steps with
// sim.time.d_t, or — with sim.time.enable_real_time in the config or --real-time on the command line — with the
// wall-clock time that passed since the previous loop iteration times sim.time.real_time_multiplier, the first step
// with d_t = 0, exactly as main.cu:72-95 does (one sayal_step + sayal_sync per iteration: d_t changes every step, so
// there is no graph to replay; the reference also blocks once per step, fluid.cu:794).
// --dump writes raw little-endian fp32 fields (reference layout) every K steps.
// --frames writes what graphics.update(fluid, d_t) (main.cu:98) would have shown every K steps, as binary PPM:
// the RGBA frame of graphics_handler.cu:258-302 with, when visual.path_line.enable / visual.arrows.enable are set,
// the path lines and arrows drawn over it.  The frame is read back asynchronously (sayal_frame_submit /
// sayal_frame_acquire): the next K steps are enqueued before the previous frame is waited for, so stepping never
// stalls on the copy the way the reference's cudaDeviceSynchronize after every cudaMemcpyAsync does.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sayal.h"

static int die(const char* what, int code) {
  std::fprintf(stderr, "sayal_run: %s failed (%d): %s\n", what, code, sayal_last_error());
  return 1;
}

static bool dump(sayal_sim* sim, const sayal_config& c, const std::string& prefix, int step) {
  static const struct { int id; const char* name; } fields[] = {
      {SAYAL_U, "u"}, {SAYAL_V, "v"}, {SAYAL_SMOKE, "smoke"}, {SAYAL_P, "p"}};
  std::vector<float> buf((size_t)c.width * c.height);
  for (auto& f : fields) {
    if (f.id == SAYAL_P && !c.enable_pressure) continue;
    if (sayal_get_field(sim, f.id, buf.data()) != SAYAL_OK) return false;
    char path[512];
    std::snprintf(path, sizeof path, "%s_%s_%06d.f32", prefix.c_str(), f.name, step);
    FILE* fp = std::fopen(path, "wb");
    if (!fp) return false;
    std::fwrite(buf.data(), sizeof(float), buf.size(), fp);
    std::fclose(fp);
  }
  return true;
}

// Bresenham, clipped; colour = r, g, b
static void draw_line(std::vector<unsigned char>& rgb, int W, int H, int x0, int y0, int x1, int y1, const int32_t* c) {
  int dx = std::abs(x1 - x0), sx = x0 < x1 ? 1 : -1, dy = -std::abs(y1 - y0), sy = y0 < y1 ? 1 : -1, err = dx + dy;
  for (int guard = 0; guard < 4 * (W + H); guard++) {
    if (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H) {
      unsigned char* px = &rgb[3 * ((size_t)y0 * W + x0)];
      px[0] = (unsigned char)c[0]; px[1] = (unsigned char)c[1]; px[2] = (unsigned char)c[2];
    }
    if (x0 == x1 && y0 == y1) break;
    int e2 = 2 * err;
    if (e2 >= dy) { err += dy; x0 += sx; }
    if (e2 <= dx) { err += dx; y0 += sy; }
  }
}

static bool write_frame(sayal_sim* sim, const sayal_config& c, const sayal_visual& vis, const uint32_t* pixels,
                        const std::string& prefix, long long step) {
  const int W = c.width, H = c.height;
  std::vector<unsigned char> rgb((size_t)W * H * 3);
  for (size_t k = 0; k < (size_t)W * H; k++) {
    rgb[3 * k] = pixels[k] >> 24; rgb[3 * k + 1] = (pixels[k] >> 16) & 255; rgb[3 * k + 2] = (pixels[k] >> 8) & 255;
  }
  if (vis.path_line_enable) {  // update_traces (graphics_handler.cu:404-421)
    int nx = 0, ny = 0;
    sayal_path_lines(sim, &vis, c.d_t, nullptr, nullptr, 0, &nx, &ny);
    const int len = vis.path_line_length;
    std::vector<int32_t> xs((size_t)nx * ny * len), ys(xs.size());
    if (!xs.empty() && sayal_path_lines(sim, &vis, c.d_t, xs.data(), ys.data(), (int32_t)xs.size(), &nx, &ny) != SAYAL_OK) return false;
    for (size_t l = 0; l < (size_t)nx * ny; l++) {
      if (xs[l * len] < 0) continue;
      for (int k = 1; k < len; k++)
        draw_line(rgb, W, H, xs[l * len + k - 1], ys[l * len + k - 1], xs[l * len + k], ys[l * len + k], vis.path_line_color);
    }
  }
  if (vis.arrows_enable) {  // update_center_velocity_arrow + draw_arrow (graphics_handler.cu:202-212, 337-356)
    int nx = 0, ny = 0;
    sayal_arrows(sim, &vis, nullptr, 0, &nx, &ny);
    std::vector<sayal_arrow> ar((size_t)nx * ny);
    if (!ar.empty() && sayal_arrows(sim, &vis, ar.data(), (int32_t)ar.size(), &nx, &ny) != SAYAL_OK) return false;
    for (const sayal_arrow& a : ar) {
      if (!a.valid) continue;
      draw_line(rgb, W, H, a.start_x, a.start_y, a.end_x, a.end_y, vis.arrows_color);
      draw_line(rgb, W, H, a.end_x, a.end_y, a.left_head_end_x, a.left_head_end_y, vis.arrows_color);
      draw_line(rgb, W, H, a.end_x, a.end_y, a.right_head_end_x, a.right_head_end_y, vis.arrows_color);
    }
  }
  char path[512];
  std::snprintf(path, sizeof path, "%s_%06lld.ppm", prefix.c_str(), step);
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return false;
  std::fprintf(fp, "P6\n%d %d\n255\n", W, H);
  std::fwrite(rgb.data(), 1, rgb.size(), fp);
  std::fclose(fp);
  return true;
}

int main(int argc, char** argv) {
  std::string config_path = "OpenSayal.conf.json", dump_prefix, frames_prefix;
  int steps = 100, device = 0, every = 0, temporal_block = -1;
  bool plain = false, no_graph = false, real_time = false;
  for (int k = 1; k < argc; k++) {
    auto need = [&](const char* flag) -> const char* {
      if (k + 1 >= argc) {
        std::fprintf(stderr, "sayal_run: %s needs a value\n", flag);
        std::exit(1);
      }
      return argv[++k];
    };
    if (!std::strcmp(argv[k], "--config")) config_path = need("--config");
    else if (!std::strcmp(argv[k], "--steps")) steps = std::atoi(need("--steps"));
    else if (!std::strcmp(argv[k], "--device")) device = std::atoi(need("--device"));
    else if (!std::strcmp(argv[k], "--dump")) dump_prefix = need("--dump");
    else if (!std::strcmp(argv[k], "--frames")) frames_prefix = need("--frames");
    else if (!std::strcmp(argv[k], "--every")) every = std::atoi(need("--every"));
    else if (!std::strcmp(argv[k], "--temporal-block")) temporal_block = std::atoi(need("--temporal-block"));
    else if (!std::strcmp(argv[k], "--plain")) plain = true;
    else if (!std::strcmp(argv[k], "--no-graph")) no_graph = true;
    else if (!std::strcmp(argv[k], "--real-time")) real_time = true;
    else {
      std::fprintf(stderr,
                   "usage: sayal_run [--config FILE] [--steps N] [--device D] [--dump PREFIX] [--frames PREFIX] [--every K] "
                   "[--plain] [--temporal-block T] [--no-graph] [--real-time]\n");
      return 1;  // main.cu:19-25 exits 1 on bad argv
    }
  }
  sayal_config cfg;
  int r = sayal_config_load(config_path.c_str(), &cfg);
  if (r != SAYAL_OK) return die("sayal_config_load", r);
  sayal_visual vis;
  r = sayal_visual_load(config_path.c_str(), &vis);
  if (r != SAYAL_OK) return die("sayal_visual_load", r);
  sayal_sim* sim = nullptr;
  r = sayal_create(&cfg, device, &sim);
  if (r != SAYAL_OK) return die("sayal_create", r);
  if (plain) sayal_set_option(sim, "projection_kernel", 0);
  if (temporal_block >= 0) sayal_set_option(sim, "temporal_block", temporal_block);
  if (no_graph) sayal_set_option(sim, "use_graph", 0);

  auto t0 = std::chrono::steady_clock::now();
  int done = 0;
  const bool outputs = !dump_prefix.empty() || !frames_prefix.empty();
  const bool overlays = vis.path_line_enable || vis.arrows_enable;  // need the state of the frame's own step
  int frames_in_flight = 0;
  auto collect_frame = [&]() -> int {  // wait for the oldest submitted frame and write it
    const uint32_t* px = nullptr;
    int64_t at = 0;
    int rr = sayal_frame_acquire(sim, &px, &at);
    if (rr != SAYAL_OK) return rr;
    frames_in_flight--;
    return write_frame(sim, cfg, vis, px, frames_prefix, (long long)at) ? SAYAL_OK : SAYAL_EIO;
  };
  real_time = real_time || cfg.enable_real_time != 0;
  bool have_prev = false;
  std::chrono::steady_clock::time_point prev_time;
  double simulated = 0.0;
  while (done < steps) {
    int chunk = (every > 0 && outputs) ? std::min(every, steps - done) : steps - done;
    if (real_time) {  // main.cu:72-95: d_t = time since the previous iteration x multiplier (0 on the first one)
      const sayal_source idle = {0, 0.f, 0.f, 0, 0};
      for (int k = 0; k < chunk; k++) {
        const auto now = std::chrono::steady_clock::now();
        if (!have_prev) prev_time = now, have_prev = true;
        const long long ns = std::chrono::duration_cast<std::chrono::nanoseconds>(now - prev_time).count();
        const float d_t = (ns != 0 ? (float)ns / 1000000000 : 0.f) * cfg.real_time_multiplier;
        r = sayal_step(sim, &idle, d_t);
        if (r == SAYAL_OK) r = sayal_sync(sim);
        if (r != SAYAL_OK) return die("sayal_step", r);
        simulated += d_t;
        prev_time = now;
      }
    } else {
      r = sayal_run(sim, chunk, cfg.d_t);
      if (r != SAYAL_OK) return die("sayal_run", r);
      simulated += (double)chunk * cfg.d_t;
    }
    done += chunk;
    if (!dump_prefix.empty() && !dump(sim, cfg, dump_prefix, done)) return die("dump", -1);
    if (!frames_prefix.empty()) {
      // the previous frame was copied while this chunk was being enqueued / executed: collect it now, then submit
      // the new one and go straight on to the next chunk
      if (frames_in_flight > 0 && (r = collect_frame()) != SAYAL_OK) return die("frame", r);
      if ((r = sayal_frame_submit(sim)) != SAYAL_OK) return die("sayal_frame_submit", r);
      frames_in_flight++;
      if (overlays && (r = collect_frame()) != SAYAL_OK) return die("frame", r);
    }
  }
  while (frames_in_flight > 0)
    if ((r = collect_frame()) != SAYAL_OK) return die("frame", r);
  r = sayal_sync(sim);
  if (r != SAYAL_OK) return die("sayal_sync", r);
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  float mn = 0, mx = 0;
  if (cfg.enable_pressure) sayal_pressure_range(sim, &mn, &mx);
  std::printf("{\"width\": %d, \"height\": %d, \"steps\": %d, \"seconds\": %.6f, \"simulated_seconds\": %.6f, \"cell_steps_per_s\": %.4e, "
              "\"kernel_launches\": %lld, \"min_pressure\": %g, \"max_pressure\": %g}\n",
              cfg.width, cfg.height, steps, sec, simulated, (double)cfg.width * cfg.height * steps / sec,
              (long long)sayal_launch_count(sim), mn, mx);
  sayal_destroy(sim);
  return 0;
}
