// `btrc <config.json>` -- command-line driver with the reference CLI's contract (src/cli/main.cpp:83-105):
// one argument (the scene file), exit code 0 on success or usage, -1 with the chained error messages on failure.
// It goes through the C ABI only.
//
// `btrc --ranks N <config.json>` renders the same scene on N GPUs of this box (new work: the reference is single-GPU,
// main.cpp:23): the driver starts one process per GPU (`--rank r`, internal), every rank loads the scene, renders its share
// (btrc_render_part: sample ranges, or interleaved image tiles when the renderer has "tile": true), the films are summed onto
// rank 0 with one ncclReduce over NVLink (btrc_reduce_films) and rank 0 resolves and runs the post-processors.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

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

static int usage() {
    std::cout << "usage: btrc config.json\n       btrc --ranks N config.json   (N GPUs of this box, one process per GPU)" << std::endl;
    return 0;
}

// the 128-byte ncclUniqueId travels from rank 0 to its siblings through a file next to nothing else: written whole, then renamed
static bool read_id(const std::string &path, unsigned char id[128]) {
    for (int tries = 0; tries < 6000; ++tries) {  // up to 60 s: rank 0 may still be initialising its device
        std::ifstream f(path, std::ios::binary);
        if (f && f.read(reinterpret_cast<char *>(id), 128)) return true;
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    return false;
}

static int run(const char *scene_path, int ranks, int rank, const std::string &id_file) {
    const bool root = rank == 0;
    if (ranks > 1 && btrc_set_device(rank)) return fail();
    btrc_comm_t *comm = nullptr;
    if (ranks > 1) {
        unsigned char id[128];
        if (root) {
            if (btrc_nccl_unique_id(id)) return fail();
            const std::string tmp = id_file + ".tmp";
            { std::ofstream f(tmp, std::ios::binary); f.write(reinterpret_cast<const char *>(id), 128); }
            std::rename(tmp.c_str(), id_file.c_str());
        } else if (!read_id(id_file, id)) {
            std::cerr << "rank " << rank << ": no NCCL id from rank 0" << std::endl;
            return -1;
        }
        if (btrc_comm_create(ranks, rank, id, &comm)) return fail();
    }
    if (root) std::cout << "parse scene / create scene / commit objects" << std::endl;
    btrc_scene_t *scene = nullptr;
    if (btrc_scene_load_json(scene_path, &scene)) return fail();
    btrc_scene_info_t info;
    if (btrc_scene_info_get(scene, &info)) return fail();
    if (root) {
        std::cout << info.width << "x" << info.height << ", " << info.spp << " spp, " << info.n_triangles << " triangles, "
                  << info.n_instances << " entities, " << info.n_lights << " lights; bvh " << info.bvh_build_seconds << " s";
        if (ranks > 1) std::cout << "; " << ranks << " GPUs";
        std::cout << std::endl << "render image" << std::endl;
    }
    const size_t n = 4 * (size_t)info.width * info.height;
    std::vector<float> color(root ? n : 0), albedo(root && info.albedo ? n : 0), normal(root && info.normal ? n : 0);
    btrc_render_stats_t st;
    if (root && btrc_set_reporter(scene, console_progress, nullptr)) return fail();  // main.cpp:62
    const auto t0 = std::chrono::steady_clock::now();
    if (ranks == 1) {
        if (btrc_render(scene, BTRC_HOST, color.data(), info.albedo ? albedo.data() : nullptr, info.normal ? normal.data() : nullptr, &st)) return fail();
    } else {
        if (btrc_film_clear(scene)) return fail();
        if (btrc_render_part(scene, ranks, rank, 0, &st)) return fail();
        if (btrc_reduce_films(scene, comm, 0)) return fail();
        if (root && btrc_resolve(scene, BTRC_HOST, color.data(), info.albedo ? albedo.data() : nullptr, info.normal ? normal.data() : nullptr)) return fail();
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%s%.3f s (device %.1f ms): %.2f Msamples/s, %.2f Mrays/s, %llu iterations\n", ranks > 1 ? ("rank " + std::to_string(rank) + ": ").c_str() : "",
                sec, st.render_ms, st.n_gen / (st.render_ms * 1e3), (st.n_closest + st.n_shadow) / (st.render_ms * 1e3), (unsigned long long)st.n_iter);
    if (root) {
        std::cout << "execute post processors" << std::endl;
        if (btrc_run_post_processors(scene, BTRC_HOST, color.data(), info.albedo ? albedo.data() : nullptr, info.normal ? normal.data() : nullptr))
            return fail();
    }
    btrc_scene_free(scene);
    btrc_comm_destroy(comm);
    return 0;
}

int main(int argc, char *argv[]) {
    int ranks = 1, rank = -1;
    std::string id_file;
    std::vector<const char *> rest;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--ranks") && i + 1 < argc) ranks = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--rank") && i + 1 < argc) rank = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--id-file") && i + 1 < argc) id_file = argv[++i];
        else rest.push_back(argv[i]);
    }
    if (rank < 0) std::cout << ">>> Btrc Renderer (btrc_b200) <<<" << std::endl;
    if (rest.size() != 1 || ranks < 1) return usage();
    if (ranks == 1) return run(rest[0], 1, 0, "");
    if (rank >= 0) return run(rest[0], ranks, rank, id_file);  // one of the per-GPU processes

    // driver: one process per GPU; no CUDA call is made in this process
    char tmpl[] = "/tmp/btrc_nccl_id_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd >= 0) close(fd);
    unlink(tmpl);
    std::vector<pid_t> pids;
    for (int r = 0; r < ranks; ++r) {
        const pid_t pid = fork();
        if (pid == 0) {
            const std::string rs = std::to_string(r), ns = std::to_string(ranks);
            execl(argv[0], argv[0], "--ranks", ns.c_str(), "--rank", rs.c_str(), "--id-file", tmpl, rest[0], (char *)nullptr);
            std::perror("btrc: exec");
            _exit(127);
        }
        if (pid < 0) { std::perror("btrc: fork"); return -1; }
        pids.push_back(pid);
    }
    int rc = 0;
    for (pid_t pid : pids) {
        int status = 0;
        waitpid(pid, &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = -1;
    }
    unlink(tmpl);
    return rc;
}
