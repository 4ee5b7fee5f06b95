// `btrc <config.json>` -- command-line driver with the reference CLI's contract (src/cli/main.cpp:83-105):
// one argument (the scene file), exit code 0 on success or usage, -1 with the chained error messages on failure.
// It goes through the C ABI only.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/btrc_b200.h"

// ConsoleReporter (builtin/reporter/console.cpp:5-33): an 80-column '=' bar redrawn whenever the percentage advanced by >= 1
static int console_progress(float percentage, void *) {
    const int width = 80, filled = (int)(percentage / 100.0f * width);
    std::string bar(filled, '=');
    bar.resize(width, ' ');
    std::printf("\r[%s] %5.1f%%", bar.c_str(), percentage);
    if (percentage >= 100.0f) std::printf("\n");
    std::fflush(stdout);
    return 0;
}

static int fail() {
    std::cerr << btrc_last_error() << std::endl;
    return -1;
}

int main(int argc, char *argv[]) {
    std::cout << ">>> Btrc Renderer (btrc_b200) <<<" << std::endl;
    if (argc != 2) {
        std::cout << "usage: btrc config.json" << std::endl;
        return 0;
    }
    std::cout << "parse scene / create scene / commit objects" << std::endl;
    btrc_scene_t *scene = nullptr;
    if (btrc_scene_load_json(argv[1], &scene)) return fail();
    btrc_scene_info_t info;
    if (btrc_scene_info_get(scene, &info)) return fail();
    std::cout << info.width << "x" << info.height << ", " << info.spp << " spp, " << info.n_triangles << " triangles, "
              << info.n_instances << " entities, " << info.n_lights << " lights; bvh " << info.bvh_build_seconds << " s" << std::endl;
    std::cout << "render image" << std::endl;
    const size_t n = 4 * (size_t)info.width * info.height;
    std::vector<float> color(n), albedo(info.albedo ? n : 0), normal(info.normal ? n : 0);
    btrc_render_stats_t st;
    if (btrc_set_reporter(scene, console_progress, nullptr)) return fail();  // main.cpp:62
    const auto t0 = std::chrono::steady_clock::now();
    if (btrc_render(scene, BTRC_HOST, color.data(), info.albedo ? albedo.data() : nullptr, info.normal ? normal.data() : nullptr, &st)) return fail();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%.3f s (device %.1f ms): %.2f Msamples/s, %.2f Mrays/s, %llu iterations\n", sec, st.render_ms,
                st.n_gen / (st.render_ms * 1e3), (st.n_closest + st.n_shadow) / (st.render_ms * 1e3), (unsigned long long)st.n_iter);
    std::cout << "execute post processors" << std::endl;
    if (btrc_run_post_processors(scene, BTRC_HOST, color.data(), info.albedo ? albedo.data() : nullptr, info.normal ? normal.data() : nullptr))
        return fail();
    btrc_scene_free(scene);
    return 0;
}
