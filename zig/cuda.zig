// cuda.zig -- Zig-side binding of the B200 back-end (include/turbomesh_gpu.h) for turbomesh's src/core.
//
// Drop this file next to src/core/smoothing/smooth.zig as src/core/smoothing/cuda.zig (see INTEGRATION.md).
// NOTE: written without a Zig toolchain at hand (none exists in the build image); it mirrors the way the
// reference binds its other foreign back-ends: `pub extern fn ... callconv(.c)` declarations as in
// src/core/smoothing/petsc_import.zig:3021 and a thin solver-like wrapper as in src/core/smoothing/umfpack.zig:14-55.
const std = @import("std");
const discrete = @import("../discrete.zig");
const boundary = @import("../boundary.zig");
const types = @import("../types.zig");
const wall_control_function = @import("wall_control_function.zig");

const log = std.log.scoped(.cuda_backend);

// ---- C ABI (include/turbomesh_gpu.h) ----------------------------------------------------------------
pub const tm_block = extern struct { ni: u64, nj: u64, xy: ?[*]f64 };
pub const tm_range = extern struct { block: u64, side: u32, _pad: u32 = 0, start: u64, end: u64 };
pub const tm_connection = extern struct { ranges: [2]tm_range, has_periodicity: i32, _pad: i32 = 0, periodicity: [2]f64 };
pub const tm_condition = extern struct { range: tm_range, kind: u32, _pad: u32 = 0 };
pub const tm_smooth_options = extern struct {
    struct_size: u32,
    solver: u32, // 0 = picard_bicgstab, 1 = relax, 2 = multigrid (FAS V-cycles, any multi-block topology, Laplace control function)
    iterations: u64,
    control_function: u32, // 0 = laplace, 1 = white
    fail_on_no_convergence: u32,
    white_ds_target: f64,
    white_theta_target: f64,
    rtol: f64,
    atol: f64,
    max_inner_iterations: u64,
    omega: f64,
    sweeps_per_iteration: u64,
    stop_max_update: f64,
    device: i32,
    inner_refinement_cycles: i32 = 0,
};
pub const tm_smooth_stats = extern struct {
    outer_iterations: u64,
    inner_iterations: u64,
    operator_applications: u64,
    nodes: u64,
    last_sumsq_x: f64,
    last_sumsq_y: f64,
    last_residual: f64,
    last_max_update: f64,
    last_inner_residual: f64,
    gpu_seconds: f64,
    converged: i32,
    streamed_chunks: i32 = 0,
};

pub extern fn tm_tfi_block(ni: u64, nj: u64, x_i_min: [*]const f64, x_i_max: [*]const f64, x_j_min: [*]const f64, x_j_max: [*]const f64, s1: [*]const f64, s2: [*]const f64, t1: [*]const f64, t2: [*]const f64, out_xy: [*]f64) callconv(.c) c_int;
pub extern fn tm_smooth_mesh(blocks: [*]tm_block, n_blocks: usize, connections: ?[*]const tm_connection, n_connections: usize, conditions: ?[*]const tm_condition, n_conditions: usize, opts: *const tm_smooth_options, stats: ?*tm_smooth_stats) callconv(.c) c_int;
pub extern fn tm_smooth_options_default(opts: *tm_smooth_options) callconv(.c) void;
pub extern fn tm_last_error() callconv(.c) [*:0]const u8;

// Device-resident handle (viewers, repeated smoothing, structured output without an AoS round trip); see the header.
pub const tm_mesh = opaque {};
pub extern fn tm_mesh_create(blocks: [*]const tm_block, n_blocks: usize, connections: ?[*]const tm_connection, n_connections: usize, conditions: ?[*]const tm_condition, n_conditions: usize, device: c_int, stream: ?*anyopaque, out: *?*tm_mesh) callconv(.c) c_int;
pub extern fn tm_mesh_destroy(mesh: ?*tm_mesh) callconv(.c) void;
pub extern fn tm_mesh_begin_smoothing(mesh: *tm_mesh, opts: *const tm_smooth_options) callconv(.c) c_int;
pub extern fn tm_mesh_smooth(mesh: *tm_mesh, opts: *const tm_smooth_options, stats: ?*tm_smooth_stats) callconv(.c) c_int;
pub extern fn tm_mesh_download_block(mesh: *tm_mesh, block: usize, xy: [*]f64) callconv(.c) c_int;
pub extern fn tm_mesh_download_block_async(mesh: *tm_mesh, block: usize, xy: [*]f64) callconv(.c) c_int;
pub extern fn tm_mesh_download_wait(mesh: *tm_mesh) callconv(.c) c_int;
/// cgns.zig:69-101 / 110-161 on the device: x[j*ni + i], y[j*ni + i] (field 0 = coordinates, 1 = control function P,Q)
pub extern fn tm_mesh_download_block_soa(mesh: *tm_mesh, block: usize, field: c_int, x: [*]f64, y: [*]f64) callconv(.c) c_int;
pub extern fn tm_release_cached_memory() callconv(.c) void;
// independent systems (one per cut of a batch), structured writer, the edge operations of the blocking
pub const tm_component_stats = extern struct { nodes: u64, iterations: [2]u64, tolerance: [2]f64, norm_b: [2]f64, norm_r: [2]f64, status: [2]i32, operator_applications: u64, restarts: u64 };
pub extern fn tm_mesh_component_count(mesh: *const tm_mesh) callconv(.c) u64;
pub extern fn tm_mesh_component_of_block(mesh: *const tm_mesh, block: usize, component: *u64) callconv(.c) c_int;
pub extern fn tm_mesh_component_stats(mesh: *const tm_mesh, component: usize, out: *tm_component_stats) callconv(.c) c_int;
pub extern fn tm_mesh_write_plot3d(mesh: *tm_mesh, grid_path: [*:0]const u8, function_path: ?[*:0]const u8) callconv(.c) c_int;
pub const tm_spline_fit_job = extern struct { n_points: u64, points: [*]const f64, n_samples: u64, params: [*]f64, second_derivs_x: [*]f64, second_derivs_y: [*]f64, sample_arc: [*]f64, total_length: *f64 };
pub extern fn tm_splines_fit(jobs: [*]const tm_spline_fit_job, n_jobs: usize, device: c_int) callconv(.c) c_int;
pub const tm_edge_view = extern struct { points: [*]const f64, clustering: [*]const f64, n: u64, start: u64, end: u64 };
pub const tm_combine_job = extern struct { views: [*]const tm_edge_view, n_views: u64, points: [*]f64, clustering: [*]f64 };
pub const tm_project_job = extern struct { points: [*]const f64, n: u64, distance: f64, out: [*]f64 };
pub extern fn tm_edges_combine(jobs: [*]const tm_combine_job, n_jobs: usize, device: c_int) callconv(.c) c_int;
pub extern fn tm_edges_project_normal(jobs: [*]const tm_project_job, n_jobs: usize, device: c_int) callconv(.c) c_int;
/// Host-only: how tm_smooth_mesh would stream a large single block (n_chunks = 0: resident); see the header.
pub extern fn tm_smooth_stream_plan(ni: u64, nj: u64, sweeps: u64, n_chunks: *u64, window_rows: *u64, window_first: ?[*]u64, owned_first: ?[*]u64) callconv(.c) c_int;
/// CUDA runtime: page-locks a host range.  tm_smooth_mesh overlaps its host<->device copies with the sweeps only when
/// the block lives in page-locked memory; a caller that smooths a large block registers `block.points.data` once
/// (flags = 0) and unregisters it before freeing.  Pageable memory is still correct, just serial.
pub extern fn cudaHostRegister(ptr: *anyopaque, size: usize, flags: c_uint) callconv(.c) c_int;
pub extern fn cudaHostUnregister(ptr: *anyopaque) callconv(.c) c_int;

pub const Error = error{ CudaBackendFailed, CudaNoDevice, CudaBadTopology, CudaNotConverged };

fn check(rc: c_int) Error!void {
    if (rc == 0) return;
    log.err("turbomesh_gpu: {s}", .{std.mem.span(tm_last_error())});
    return switch (rc) {
        -2 => error.CudaNoDevice,
        -4, -5 => error.CudaBadTopology,
        -6 => error.CudaNotConverged,
        else => error.CudaBackendFailed,
    };
}

/// Options of the `"cuda"` variant of `solver.Option` (JSON: `"solver": {"cuda": {"method": "picard_bicgstab"}}`).
pub const Option = struct {
    method: enum { picard_bicgstab, relax, multigrid } = .picard_bicgstab, // = tm_solver (0, 1, 2)
    rtol: f64 = 1e-6, // BiCGStab.zig:20
    atol: f64 = 1e-8, // BiCGStab.zig:21
    max_inner_iterations: u64 = 1000, // BiCGStab.zig:19
    omega: f64 = 1.0,
    sweeps_per_iteration: u64 = 1,
    stop_max_update: f64 = 0.0, // > 0: stop the outer loop once the mesh moves by less than this (relax / multigrid)
    device: i32 = -1,
};

/// Replacement for the body of `tfi.linear2dBoundaryBlendedControlFunction` as called from
/// `discrete.Block2d.init` (discrete.zig:147-158): same arguments, result written into `data`.
pub fn tfi(
    data: *types.Mat2d,
    x_i_min: []const types.Vec2d,
    x_i_max: []const types.Vec2d,
    x_j_min: []const types.Vec2d,
    x_j_max: []const types.Vec2d,
    s1: []const types.Float,
    s2: []const types.Float,
    t1: []const types.Float,
    t2: []const types.Float,
) Error!void {
    // Vec2d is `struct { data: [2]f64 }` (types.zig:16-27): a slice of it is the interleaved x,y array the ABI expects
    try check(tm_tfi_block(
        data.size[0],
        data.size[1],
        @ptrCast(x_i_min.ptr),
        @ptrCast(x_i_max.ptr),
        @ptrCast(x_j_min.ptr),
        @ptrCast(x_j_max.ptr),
        s1.ptr,
        s2.ptr,
        t1.ptr,
        t2.ptr,
        @ptrCast(data.data.ptr),
    ));
}

/// Replacement for `smooth.mesh` (smooth.zig:74-166) when `solver_option == .cuda`: smooths all blocks in place.
pub fn mesh(
    allocator: std.mem.Allocator,
    mesh_data: *discrete.Mesh,
    iterations: usize,
    option: Option,
    control_function_algorithm: wall_control_function.Algorithm,
) !void {
    const blocks = try allocator.alloc(tm_block, mesh_data.blocks.items.len);
    defer allocator.free(blocks);
    for (mesh_data.blocks.items, blocks) |b, *out| {
        out.* = .{ .ni = b.points.size[0], .nj = b.points.size[1], .xy = @ptrCast(b.points.data.ptr) };
    }

    const connections = try allocator.alloc(tm_connection, mesh_data.connections.items.len);
    defer allocator.free(connections);
    for (mesh_data.connections.items, connections) |c, *out| {
        out.* = .{
            .ranges = .{ toRange(c.ranges[0]), toRange(c.ranges[1]) },
            .has_periodicity = if (c.periodicity != null) 1 else 0,
            .periodicity = if (c.periodicity) |p| p.data else .{ 0, 0 },
        };
    }

    const conditions = try allocator.alloc(tm_condition, mesh_data.boundary_conditions.items.len);
    defer allocator.free(conditions);
    for (mesh_data.boundary_conditions.items, conditions) |bc, *out| {
        out.* = .{ .range = toRange(bc.range), .kind = @intFromEnum(bc.kind) }; // wall = 0, inlet = 1, outlet = 2 (boundary.zig:172-176)
    }

    var opts: tm_smooth_options = undefined;
    tm_smooth_options_default(&opts);
    opts.solver = @intFromEnum(option.method);
    opts.iterations = iterations;
    opts.rtol = option.rtol;
    opts.atol = option.atol;
    opts.max_inner_iterations = option.max_inner_iterations;
    opts.omega = option.omega;
    opts.sweeps_per_iteration = option.sweeps_per_iteration;
    opts.stop_max_update = option.stop_max_update;
    opts.device = option.device;
    switch (control_function_algorithm) {
        .laplace => opts.control_function = 0,
        .white => |w| {
            opts.control_function = 1;
            opts.white_ds_target = w.ds_target;
            opts.white_theta_target = w.theta_target;
        },
    }

    var stats: tm_smooth_stats = undefined;
    try check(tm_smooth_mesh(blocks.ptr, blocks.len, connections.ptr, connections.len, conditions.ptr, conditions.len, &opts, &stats));
    log.info("\tresidual: {} ({} outer, {} inner iterations, {d:.3} s on the GPU)", .{ stats.last_residual, stats.outer_iterations, stats.inner_iterations, stats.gpu_seconds });
}

fn toRange(r: boundary.Range) tm_range {
    // boundary.Side is declared i_min, i_max, j_min, j_max (boundary.zig:8-13) = 0..3, the ABI's tm_side
    return .{ .block = r.block, .side = @intFromEnum(r.side), .start = r.start, .end = r.end };
}
