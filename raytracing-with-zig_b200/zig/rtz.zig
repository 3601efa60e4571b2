//! rtz.zig — the Zig side of the drop-in: `extern` view of include/rtz.h plus the two helpers
//! `Camera.render` needs.  This file and the three small edits in INTEGRATION.md are everything a
//! maintainer of raytracing-with-zig adds; Scene / CameraBuilder / main stay exactly as they are.
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no zig toolchain.  It is kept
//! deliberately thin (no logic beyond field copies) and mirrors, line for line, the C++ host in
//! ../host/rtz_host.hpp (`Camera::flat`, `HittableList::flat`, `Camera::render`), which IS compiled
//! and tested against the oracle.
const std = @import("std");

// ---- include/rtz.h, as extern structs (fixed-width ints and f64 only) -------------------------
pub const RTZ_ABI_VERSION: i32 = 2;
pub const RTZ_OK: i32 = 0;
pub const RTZ_MODE_PATH: i32 = 0;

pub const rtz_sphere = extern struct {
    center: [3]f64,
    radius: f64,
    mat_type: i32, // 0 lambertian, 1 metal, 2 dielectric  (material.zig MaterialType order)
    reserved: i32 = 0,
    albedo: [3]f64,
    fuzz: f64,
    refraction_index: f64,
};

pub const rtz_camera = extern struct {
    width: u64,
    height: u64,
    center: [3]f64,
    pixel0: [3]f64,
    du: [3]f64,
    dv: [3]f64,
    defocus_disk_u: [3]f64,
    defocus_disk_v: [3]f64,
    defocus_angle: f64,
    samples_per_pixel: u64,
    bounce_max: u64,
    pixel_samples_scale: f64,
    t_min: f64,
    t_max: f64,
    seed: u64,
    has_seed: i32,
    mode: i32,
};

pub const rtz_stats = extern struct {
    samples: u64,
    segments: u64,
    sphere_tests: u64,
    depth_capped: u64,
    absorbed: u64,
    kernel_launches: u64,
    trace_ms: f64,
    resolve_ms: f64,
    total_ms: f64,
    seed_used: u64,
    nan_samples: u64,
    gpus: u32,
    gather: u32,
    gather_ms: f64,
};

// Layout contract with include/rtz.h, checked by the Zig compiler when this file is built AND by this
// repository's CI without a Zig compiler: tests/test_abi_symbols.py parses the numbers below and compares them
// with sizeof / offsetof from a C compiler on rtz.h, and the field lists above with the header's.
comptime {
    std.debug.assert(@sizeOf(rtz_sphere) == 80);
    std.debug.assert(@offsetOf(rtz_sphere, "center") == 0);
    std.debug.assert(@offsetOf(rtz_sphere, "radius") == 24);
    std.debug.assert(@offsetOf(rtz_sphere, "mat_type") == 32);
    std.debug.assert(@offsetOf(rtz_sphere, "reserved") == 36);
    std.debug.assert(@offsetOf(rtz_sphere, "albedo") == 40);
    std.debug.assert(@offsetOf(rtz_sphere, "fuzz") == 64);
    std.debug.assert(@offsetOf(rtz_sphere, "refraction_index") == 72);
    std.debug.assert(@sizeOf(rtz_camera) == 224);
    std.debug.assert(@offsetOf(rtz_camera, "width") == 0);
    std.debug.assert(@offsetOf(rtz_camera, "height") == 8);
    std.debug.assert(@offsetOf(rtz_camera, "center") == 16);
    std.debug.assert(@offsetOf(rtz_camera, "pixel0") == 40);
    std.debug.assert(@offsetOf(rtz_camera, "du") == 64);
    std.debug.assert(@offsetOf(rtz_camera, "dv") == 88);
    std.debug.assert(@offsetOf(rtz_camera, "defocus_disk_u") == 112);
    std.debug.assert(@offsetOf(rtz_camera, "defocus_disk_v") == 136);
    std.debug.assert(@offsetOf(rtz_camera, "defocus_angle") == 160);
    std.debug.assert(@offsetOf(rtz_camera, "samples_per_pixel") == 168);
    std.debug.assert(@offsetOf(rtz_camera, "bounce_max") == 176);
    std.debug.assert(@offsetOf(rtz_camera, "pixel_samples_scale") == 184);
    std.debug.assert(@offsetOf(rtz_camera, "t_min") == 192);
    std.debug.assert(@offsetOf(rtz_camera, "t_max") == 200);
    std.debug.assert(@offsetOf(rtz_camera, "seed") == 208);
    std.debug.assert(@offsetOf(rtz_camera, "has_seed") == 216);
    std.debug.assert(@offsetOf(rtz_camera, "mode") == 220);
    std.debug.assert(@sizeOf(rtz_stats) == 104);
    std.debug.assert(@offsetOf(rtz_stats, "samples") == 0);
    std.debug.assert(@offsetOf(rtz_stats, "segments") == 8);
    std.debug.assert(@offsetOf(rtz_stats, "sphere_tests") == 16);
    std.debug.assert(@offsetOf(rtz_stats, "depth_capped") == 24);
    std.debug.assert(@offsetOf(rtz_stats, "absorbed") == 32);
    std.debug.assert(@offsetOf(rtz_stats, "kernel_launches") == 40);
    std.debug.assert(@offsetOf(rtz_stats, "trace_ms") == 48);
    std.debug.assert(@offsetOf(rtz_stats, "resolve_ms") == 56);
    std.debug.assert(@offsetOf(rtz_stats, "total_ms") == 64);
    std.debug.assert(@offsetOf(rtz_stats, "seed_used") == 72);
    std.debug.assert(@offsetOf(rtz_stats, "nan_samples") == 80);
    std.debug.assert(@offsetOf(rtz_stats, "gpus") == 88);
    std.debug.assert(@offsetOf(rtz_stats, "gather") == 92);
    std.debug.assert(@offsetOf(rtz_stats, "gather_ms") == 96);
}

pub extern "rtz" fn rtz_render(camera: *const rtz_camera, spheres: [*]const rtz_sphere, n_spheres: u64, rgb_out: [*]u8, stats_out: ?*rtz_stats) i32;
/// The same call on `num_gpus` GPUs of the box (<= 0: all of them) driven by THIS process: no launcher, no MPI.
pub extern "rtz" fn rtz_render_multi(camera: *const rtz_camera, spheres: [*]const rtz_sphere, n_spheres: u64, num_gpus: i32, rgb_out: [*]u8, stats_out: ?*rtz_stats) i32;
pub extern "rtz" fn rtz_abi_version() i32;
pub extern "rtz" fn rtz_write_ppm(path: [*:0]const u8, width: u64, height: u64, rgb: [*]const u8) i32;
pub extern "rtz" fn rtz_strerror(status: i32) [*:0]const u8;
// resident API + device-side Scene.generateWorld (kind 0 final world, 1 chapter 13, 2 sweep of n spheres)
pub const rtz_context = opaque {};
pub extern "rtz" fn rtz_context_create(device: i32, stream: ?*anyopaque, ctx_out: *?*rtz_context) i32;
pub extern "rtz" fn rtz_context_destroy(ctx: *rtz_context) i32;
pub extern "rtz" fn rtz_scene_generate(ctx: *rtz_context, kind: i32, seed: u64, n_spheres: u64, spheres_out: ?[*]rtz_sphere, cap: u64, n_out: ?*u64, prng_state_out: ?*[4]u64) i32;
pub extern "rtz" fn rtz_last_error() [*:0]const u8;

pub const RenderError = error{RenderFailed};

fn v3(v: @Vector(3, f64)) [3]f64 {
    return .{ v[0], v[1], v[2] };
}

/// HittableList -> rtz_sphere[]  (hittable.zig:43-62; the list only ever holds `.sphere`).
/// `world` is the reference's `HittableList`, `Material` its tagged union (material.zig:126-129).
pub fn flattenWorld(alloc: std.mem.Allocator, world: anytype) ![]rtz_sphere {
    const out = try alloc.alloc(rtz_sphere, world.objects.items.len);
    for (world.objects.items, 0..) |item, i| {
        const s = item.sphere;
        var flat = rtz_sphere{
            .center = v3(s.center),
            .radius = s.radius,
            .mat_type = 0,
            .albedo = .{ 1, 1, 1 },
            .fuzz = 0,
            .refraction_index = 1.0,
        };
        switch (s.mat) {
            .lambertian => |l| {
                flat.mat_type = 0;
                flat.albedo = v3(l.albedo);
            },
            .metal => |m| {
                flat.mat_type = 1;
                flat.albedo = v3(m.albedo);
                flat.fuzz = m.fuzz;
            },
            .dielectric => |d| {
                flat.mat_type = 2;
                flat.refraction_index = d.refractionIndex;
            },
        }
        out[i] = flat;
    }
    return out;
}

/// The fields of `Camera` that render() reads (camera.zig:82-103), by value.
pub fn flattenCamera(cam: anytype) rtz_camera {
    return .{
        .width = cam.image.width,
        .height = cam.image.height,
        .center = v3(cam.center),
        .pixel0 = v3(cam.pixel0),
        .du = v3(cam.du),
        .dv = v3(cam.dv),
        .defocus_disk_u = v3(cam.defocusDiskU),
        .defocus_disk_v = v3(cam.defocusDiskV),
        .defocus_angle = cam.defocusAngle,
        .samples_per_pixel = cam.samplesPerPixel,
        .bounce_max = cam.bounceMax,
        .pixel_samples_scale = cam.pixelSamplesScale,
        .t_min = cam.scene.interval.min,
        .t_max = cam.scene.interval.max,
        .seed = cam.scene.seed orelse 0,
        .has_seed = if (cam.scene.seed != null) 1 else 0,
        .mode = RTZ_MODE_PATH,
    };
}

/// Body of `Camera.render` (camera.zig:123-145) on the B200: one C-ABI call instead of the
/// rows x columns x samples loop nest, then the same P6 file PPM.saveBinary writes.
/// `num_gpus` comes from the new build option -DnumGpus (1 by default; 0 = every GPU of the box).
pub fn render(cam: anytype, comptime path: [:0]const u8, num_gpus: i32) RenderError!void {
    if (rtz_abi_version() != RTZ_ABI_VERSION) return error.RenderFailed;
    const spheres = flattenWorld(cam.alloc, cam.scene.world) catch return error.RenderFailed;
    defer cam.alloc.free(spheres);
    const c = flattenCamera(cam);
    const rgb = cam.alloc.alloc(u8, 3 * cam.image.width * cam.image.height) catch return error.RenderFailed;
    defer cam.alloc.free(rgb);
    const status = if (num_gpus == 1)
        rtz_render(&c, spheres.ptr, spheres.len, rgb.ptr, null)
    else
        rtz_render_multi(&c, spheres.ptr, spheres.len, num_gpus, rgb.ptr, null);
    if (status != RTZ_OK) {
        std.log.err("rtz_render: {s}: {s}", .{ rtz_strerror(status), rtz_last_error() });
        return error.RenderFailed;
    }
    if (rtz_write_ppm(path.ptr, c.width, c.height, rgb.ptr) != RTZ_OK) return error.RenderFailed;
}
