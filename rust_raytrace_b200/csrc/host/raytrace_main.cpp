// raytrace_main.cpp — `raytrace_b200`, the main.rs-equivalent driver (raytrace/src/main.rs:88-227) over the C ABI.
//
// Builds the scene of main.rs:116-152 (dummy triangle, the teapot through parse_obj with main.rs's transform, the two
// mirror disks), the camera of main.rs:166-173, renders with the B200 caster where main.rs calls
// `caster.walk_rays(&v, &scene, &mut data, threads, false)` (main.rs:191-200), prints ProgressCtx::print_stats's lines
// (progress.rs:157-185) and writes the PNG (`write_png`, main.rs:205).  The SDL window and the DebugCtx CSV dumps of
// main.rs are not part of the path and are not reproduced.  There is no CPU rendering path: without a GPU the
// program exits with the library's error.
//
//   raytrace_b200 [--size WxH] [--maxdepth D] [--spp S] [--gpus N] [--mesh teapot.obj|teapot_mesh.bin]
//                 [--out test.png] [--rgb8] [--instanced] [--deterministic] [--light x,y,z,len2] [--seed K]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "raytrace_host.hpp"

using namespace raytrace;

static void die(const char* what) {
    std::fprintf(stderr, "raytrace_b200: %s: %s\n", what, rtb_last_error());
    std::exit(1);
}

int main(int argc, char** argv) {
    uint32_t width = 64, height = 64, maxdepth = 5, spp = 1;      // main.rs:108-110, :172-173
    int gpus = 1;
    uint64_t seed = 0;
    std::string mesh = "teapot_tri.obj", out = "test.png";        // main.rs:92, :118
    bool rgb8 = false, instanced = false, deterministic = false, have_light = false;
    float light[4] = {0, 0, 0, 0};
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "--size") { if (std::sscanf(next(), "%ux%u", &width, &height) != 2) return 2; }
        else if (a == "--maxdepth") maxdepth = (uint32_t)std::atoi(next());
        else if (a == "--spp") spp = (uint32_t)std::atoi(next());
        else if (a == "--gpus") gpus = std::atoi(next());
        else if (a == "--seed") seed = std::strtoull(next(), nullptr, 10);
        else if (a == "--mesh") mesh = next();
        else if (a == "--out") out = next();
        else if (a == "--rgb8") rgb8 = true;
        else if (a == "--instanced") instanced = true;
        else if (a == "--deterministic") deterministic = true;
        else if (a == "--light") { if (std::sscanf(next(), "%f,%f,%f,%f", light, light + 1, light + 2, light + 3) != 4) return 2; have_light = true; }
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    const float aspect = float(height) / float(width);            // main.rs:96-110

    // ---- scene, main.rs:116-152 ----
    const Color orange = make_color(252, 119, 0), grey = make_color(230, 230, 230), dark = make_color(40, 40, 40);
    const SurfaceKind pot = deterministic ? SurfaceKind::Solid(orange) : SurfaceKind::Matte(orange, 0.2f);
    const SurfaceKind d1 = SurfaceKind::Reflective(deterministic ? 0.0f : 0.0002f, grey, 0.7f);
    const SurfaceKind d2 = SurfaceKind::Reflective(deterministic ? 0.0f : 0.002f, grey, 0.7f);
    const SurfaceKind side = deterministic ? SurfaceKind::Solid(dark) : SurfaceKind::Matte(dark, 0.2f);
    const Basis tf = create_transform(make_vec(0.f, 0.3f, 1.f).unit(), to_radians(270.f));
    const Vec3 offset = make_vec(0.f, 0.5f, 5.f);
    obj_parser::Mesh m;
    std::vector<Triangle> disks, tris;
    try {
        m = mesh.size() > 4 && mesh.substr(mesh.size() - 4) == ".obj" ? obj_parser::read_obj(mesh) : obj_parser::read_mesh_bin(mesh);
        disks = make_disk(make_vec(4.f, 4.f, 7.f), make_vec(-0.3f, -0.55f, -0.5f).unit(), 2.f, 0.1f, 50, d1, side, -1.f);
        const auto dk2 = make_disk(make_vec(4.f, -3.f, 5.f), make_vec(-0.5f, 2.0f, -0.5f).unit(), 1.f, 0.04f, 50, d2, side, -1.f);
        disks.insert(disks.end(), dk2.begin(), dk2.end());
        if (!instanced) {
            tris.push_back(make_dummy_triangle());
            const auto t = obj_parser::mesh_to_triangles(m, offset, 1.0f, tf, pot, 0.05f);
            tris.insert(tris.end(), t.begin(), t.end());
            tris.insert(tris.end(), disks.begin(), disks.end());
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "raytrace_b200: %s\n", e.what());
        return 1;
    }
    const float root_orig[3] = {0.f, 0.f, 20.1f};                 // build_bounding_box(.., (0,0,20.1), 20., 10, 19), main.rs:160-164
    const float root_len2 = 20.f;

    if (rtb_init(gpus, nullptr) != RTB_OK) die("rtb_init");
    rtb_scene* scene = nullptr;
    if (instanced) {                                              // the teapot's Triangles are computed on the GPU
        RtbMeshInstance inst;
        std::memset(&inst, 0, sizeof inst);
        const Vec3 rows[3] = {tf.r0, tf.r1, tf.r2};
        for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) inst.transform_rows[3 * r + k] = rows[r].v[k];
        for (int k = 0; k < 3; ++k) { inst.offset[k] = offset.v[k]; inst.color[k] = pot.color.v[k]; }
        inst.scale = 1.0f; inst.edge_thickness = 0.05f; inst.kind = pot.kind; inst.alpha = pot.alpha; inst.scattering = pot.scattering;
        if (rtb_scene_create_instanced(m.verts.data(), (uint32_t)(m.verts.size() / 3), m.faces.data(), (uint32_t)(m.faces.size() / 3),
                                       &inst, 1, disks.data(), (uint32_t)disks.size(), root_orig, root_len2, &scene) != RTB_OK)
            die("rtb_scene_create_instanced");
    } else if (rtb_scene_create(tris.data(), (uint32_t)tris.size(), root_orig, root_len2, &scene) != RTB_OK) {
        die("rtb_scene_create");
    }
    if (have_light && rtb_scene_set_light(scene, light, light[3]) != RTB_OK) die("rtb_scene_set_light");
    RtbSceneInfo info;
    rtb_scene_info(scene, &info);

    // ---- camera, main.rs:166-173 ----
    Viewport v = create_viewport(width, height, 1.f, 1.f * aspect, make_vec(2.f, 0.f, 0.f), make_vec(0.f, 0.f, 1.f).unit(), 90.f,
                                 to_radians(0.f), maxdepth, spp);
    v.seed = seed;

    // ---- walk_rays + print_stats + write_png, main.rs:190-205 ----
    const size_t npix = size_t(width) * height;
    std::vector<float> data;
    std::vector<uint8_t> rgb;
    RtbStats st;
    std::memset(&st, 0, sizeof st);
    const auto t0 = std::chrono::steady_clock::now();
    int rc;
    if (rgb8) { rgb.assign(npix * 3, 0); rc = rtb_render_rgb8(scene, &v, rgb.data(), &st); }
    else {
        data.assign(npix * 4, 0.f);
        rc = (spp > 1 && gpus != 1) ? rtb_render_progressive(scene, &v, data.data(), &st) : rtb_render(scene, &v, data.data(), nullptr, nullptr, &st);
    }
    if (rc != RTB_OK) die("render");
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // ProgressCtx::print_stats, progress.rs:157-185
    std::printf("Processed %.3f million rays in %.3f seconds. %.3f million rays/s\n", double(st.rays) / 1e6, secs,
                double(st.rays) / secs / 1e6);
    std::printf("GPU Render: %.3f\nGPU Total: %.3f\n\n", st.ms_render * 1e-3, st.ms_total * 1e-3);
    std::printf("GPU launches: %u\nRays: %llu\nReferences: %u\nTriangles: %u\n", st.kernel_launches, (unsigned long long)st.rays,
                info.n_refs, info.n_tris);
    const bool ok = rgb8 ? write_png_rgb8(out, width, height, rgb.data()) : write_png(out, width, height, data.data());
    rtb_scene_destroy(scene);
    if (!ok) { std::fprintf(stderr, "raytrace_b200: cannot write %s\n", out.c_str()); return 1; }
    return 0;
}
