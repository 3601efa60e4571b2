// build_snippet.zig — the lines to add to the reference's build.zig (build.zig:27-36 for the
// executable, :56-68 for the test step) so both link the B200 library.  See INTEGRATION.md.
//
//     const rtz_dir = b.option([]const u8, "rtzDir", "directory holding librtz.so") orelse "raytracing-with-zig_b200/csrc";
//
//     exe.addLibraryPath(.{ .cwd_relative = rtz_dir });
//     exe.addRPath(.{ .cwd_relative = rtz_dir });
//     exe.linkSystemLibrary("rtz");
//     exe.linkLibC();
//
//     exe_unit_tests.addLibraryPath(.{ .cwd_relative = rtz_dir });
//     exe_unit_tests.addRPath(.{ .cwd_relative = rtz_dir });
//     exe_unit_tests.linkSystemLibrary("rtz");
//     exe_unit_tests.linkLibC();
