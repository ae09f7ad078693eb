// sayal_run — headless driver over the C ABI; the replacement for the SDL-bound loop of
// /root/reference/src/main.cu:42-107 (ConfigParser().parse(); Fluid fluid(config); loop { fluid.update }).
//
//   sayal_run [--config OpenSayal.conf.json] [--steps N] [--device D] [--dump PREFIX] [--every K]
//             [--plain] [--temporal-block T] [--no-graph]
//
// Like the reference it reads ./OpenSayal.conf.json by default (config_parser.cpp:13) and steps with
// sim.time.d_t (main.cu:91-95; real-time d_t needs a display loop and is not offered headless).
// --dump writes raw little-endian fp32 fields (reference layout) every K steps: the "frame readback"
// of graphics_handler.cu:287-302, decoupled from stepping.
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

int main(int argc, char** argv) {
  std::string config_path = "OpenSayal.conf.json", dump_prefix;
  int steps = 100, device = 0, every = 0, temporal_block = -1;
  bool plain = false, no_graph = false;
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
    else if (!std::strcmp(argv[k], "--every")) every = std::atoi(need("--every"));
    else if (!std::strcmp(argv[k], "--temporal-block")) temporal_block = std::atoi(need("--temporal-block"));
    else if (!std::strcmp(argv[k], "--plain")) plain = true;
    else if (!std::strcmp(argv[k], "--no-graph")) no_graph = true;
    else {
      std::fprintf(stderr,
                   "usage: sayal_run [--config FILE] [--steps N] [--device D] [--dump PREFIX] [--every K] "
                   "[--plain] [--temporal-block T] [--no-graph]\n");
      return 1;  // main.cu:19-25 exits 1 on bad argv
    }
  }
  sayal_config cfg;
  int r = sayal_config_load(config_path.c_str(), &cfg);
  if (r != SAYAL_OK) return die("sayal_config_load", r);
  if (cfg.viscosity != 0.f)
    std::fprintf(stderr, "sayal_run: note: fluid.viscosity=%g is ignored (the reference's diffusion loop is racy and "
                         "out of scope, DESIGN.md H1)\n", cfg.viscosity);
  sayal_sim* sim = nullptr;
  r = sayal_create(&cfg, device, &sim);
  if (r != SAYAL_OK) return die("sayal_create", r);
  if (plain) sayal_set_option(sim, "projection_kernel", 0);
  if (temporal_block >= 0) sayal_set_option(sim, "temporal_block", temporal_block);
  if (no_graph) sayal_set_option(sim, "use_graph", 0);

  auto t0 = std::chrono::steady_clock::now();
  int done = 0;
  while (done < steps) {
    int chunk = (every > 0 && !dump_prefix.empty()) ? std::min(every, steps - done) : steps - done;
    r = sayal_run(sim, chunk, cfg.d_t);
    if (r != SAYAL_OK) return die("sayal_run", r);
    done += chunk;
    if (!dump_prefix.empty() && !dump(sim, cfg, dump_prefix, done)) return die("dump", -1);
  }
  r = sayal_sync(sim);
  if (r != SAYAL_OK) return die("sayal_sync", r);
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  float mn = 0, mx = 0;
  if (cfg.enable_pressure) sayal_pressure_range(sim, &mn, &mx);
  std::printf("{\"width\": %d, \"height\": %d, \"steps\": %d, \"seconds\": %.6f, \"cell_steps_per_s\": %.4e, "
              "\"kernel_launches\": %lld, \"min_pressure\": %g, \"max_pressure\": %g}\n",
              cfg.width, cfg.height, steps, sec, (double)cfg.width * cfg.height * steps / sec,
              (long long)sayal_launch_count(sim), mn, mx);
  sayal_destroy(sim);
  return 0;
}
