// test_driver.cpp — the reference's tests/simulation/unit/test_io.cpp restated for this build's
// config reader and CDF-5 writer (no googletest, no PnetCDF here).  Without arguments only the
// CPU-only tests run (configuration); `--gpu` adds the snapshot round trip, which needs a device.
//   Unit_IO_Yaml.LoadsNestedBlocksAndBC          test_io.cpp:34-47
//   Unit_IO_CLI.SimpleScalarOverrides            test_io.cpp:49-72
//   Unit_IO_CLI.ICOverridesTakePrecedence        test_io.cpp:74-89
//   Unit_IO_BC.ParseRoundtrip / BcToStringDefaultCase   :91-112, :142-145
//   Unit_IO_CLI.Invalid{BoundaryCondition,GridSize,Timestep}Throws, InvalidICPresetThrows  :114-140
//   Unit_IO_Yaml.MissingBlocksStillWork          :147-160
//   Unit_IO_CLI.OverridesWithSpaceSeparator / MergedConfigNoYaml   :162-176
//   Unit_IO_File.WriteNetCDFAndReadBack / WriteMetadataAndReadBack  :178-270 (through our own reader)
#include <mpi.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "decomp.hpp"
#include "field.hpp"
#include "init.hpp"
#include "io.hpp"

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        ++g_checks;                                                     \
        if (!(cond)) {                                                  \
            ++g_fail;                                                   \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
        }                                                               \
    } while (0)
template <class F>
static bool throws_runtime(F f) {
    try {
        f();
    } catch (const std::runtime_error&) {
        return true;
    } catch (...) {
    }
    return false;
}

static const char* kDevYaml =  // the values of the reference's configs/dev.yaml, incl. its ic.params quirk
    "grid:    { nx: 512, ny: 512, dx: 1.0, dy: 1.0 }\n"
    "physics: { D: 0.05, vx: 0.5, vy: 0.0 }\n"
    "time:    { dt: 0.1, steps: 1000, out_every: 100 }\n"
    "bc:\n  left: dirichlet\n  right: neumann\n  bottom: periodic\n  top: dirichlet\n"
    "output:  { prefix: \"dev\" }\n\n"
    "ic:\n  preset: gaussian_hotspot\n  file:   \"inputs/ic_global.nc\"\n  params:\n    A: 7.0\n    sigma_frac: 0.5\n";

static void config_tests() {
    const std::string dev = "tmp_dev_cfg.yaml";
    {
        std::ofstream o(dev);
        o << kDevYaml;
    }
    {
        SimConfig cfg = load_yaml_file(dev);
        CHECK(cfg.nx == 512 && cfg.ny == 512 && cfg.dt == 0.1 && cfg.steps == 1000 && cfg.out_every == 100);
        CHECK(cfg.D == 0.05 && cfg.vx == 0.5 && cfg.vy == 0.0 && cfg.dx == 1.0);
        CHECK(bc_to_string(cfg.bc.left) == "dirichlet" && bc_to_string(cfg.bc.right) == "neumann");
        CHECK(bc_to_string(cfg.bc.bottom) == "periodic" && bc_to_string(cfg.bc.top) == "dirichlet");
        CHECK(cfg.output_prefix == "dev");
        CHECK(cfg.ic.A == 1.0 && cfg.ic.sigma_frac == 0.05);  // ic.params.* is ignored
        CHECK(cfg.ic.preset == "gaussian_hotspot");
    }
    {
        const char* tmpfile = "tmp_test.yaml";
        {
            std::ofstream ofs(tmpfile);
            ofs << "grid: { nx: 64, ny: 64, dx: 1.0, dy: 1.0 }\n"
                << "physics: { D: 0.01, vx: 0.0, vy: 0.0 }\n"
                << "time: { dt: 0.1, steps: 10, out_every: 5 }\n"
                << "bc: dirichlet\n"
                << "output: { prefix: \"from_yaml\" }\n";
        }
        std::vector<std::string> args = {"--nx=128", "--ny=256", "--dt=0.2", "--bc.left=periodic",
                                         "--output_prefix=from_cli"};
        SimConfig m = merged_config(std::string(tmpfile), args);
        CHECK(m.nx == 128 && m.ny == 256 && m.dt == 0.2);
        CHECK(bc_to_string(m.bc.left) == "periodic" && bc_to_string(m.bc.right) == "dirichlet");
        CHECK(m.output_prefix == "from_cli" && m.D == 0.01 && m.steps == 10 && m.out_every == 5);
        SimConfig y = load_yaml_file(tmpfile);
        CHECK(y.output_prefix == "from_yaml");
        std::remove(tmpfile);
    }
    {
        std::vector<std::string> args = {"--ic.mode=preset",      "--ic.preset=constant_zero", "--ic.A=999.0",
                                         "--ic.sigma_frac=0.25", "--ic.xc_frac=0.1",          "--ic.yc_frac=0.2"};
        SimConfig m = merged_config(dev, args);
        CHECK(m.ic.mode == "preset" && m.ic.preset == "constant_zero" && m.ic.A == 999.0);
        CHECK(m.ic.sigma_frac == 0.25 && m.ic.xc_frac == 0.1 && m.ic.yc_frac == 0.2);
    }
    CHECK(bc_from_string("dirichlet") == BCType::Dirichlet && bc_from_string("neumann") == BCType::Neumann &&
          bc_from_string("periodic") == BCType::Periodic);
    CHECK(bc_from_string("FIXED") == BCType::Dirichlet && bc_from_string("noflux") == BCType::Neumann &&
          bc_from_string("zero-flux") == BCType::Neumann && bc_from_string("period") == BCType::Periodic);
    CHECK(bc_to_string(static_cast<BCType>(999)) == "dirichlet");
    CHECK(throws_runtime([] { merged_config(std::nullopt, {"--bc.left=foobar"}); }));
    CHECK(throws_runtime([] { merged_config(std::nullopt, {"--nx=-10", "--ny=128"}); }));
    CHECK(throws_runtime([] { merged_config(std::nullopt, {"--dt=0.0", "--steps=10"}); }));
    CHECK(throws_runtime([] { merged_config(std::nullopt, {"--out_every=0"}); }));
    {
        SimConfig cfg;
        cfg.ic.mode = "preset";
        cfg.ic.preset = "notarealpreset";
        Field f(cfg.nx, cfg.ny, 0, cfg.dx, cfg.dy);
        Decomp2D dec;
        CHECK(throws_runtime([&] { apply_initial_condition(dec, f, cfg); }));
        cfg.ic.mode = "file";
        CHECK(throws_runtime([&] { apply_initial_condition(dec, f, cfg); }));
    }
    {
        const std::string fname = "minimal.yaml";
        {
            std::ofstream ofs(fname);
            ofs << "nx: 4\nny: 5\ndx: 1.0\ndy: 1.0\n"
                << "dt: 0.1\nsteps: 2\nout_every: 1\n";
        }
        SimConfig cfg = load_yaml_file(fname);
        CHECK(cfg.nx == 4 && cfg.ny == 5 && cfg.steps == 2 && cfg.out_every == 1);
        std::remove(fname.c_str());
    }
    {
        SimConfig m = merged_config(std::nullopt, {"--nx", "42", "--dy", "2.5", "--output.prefix", "cli_space"});
        CHECK(m.nx == 42 && m.dy == 2.5 && m.output_prefix == "cli_space");
        SimConfig c = merged_config(std::nullopt, {"--nx=8", "--ny=8", "--dt=0.1", "--steps=1"});
        CHECK(c.nx == 8 && c.ny == 8);
        // --bc=periodic is not a recognised key: silently ignored, sides stay Dirichlet
        SimConfig q = merged_config(std::nullopt, {"--bc=periodic", "--unknown=1"});
        CHECK(bc_to_string(q.bc.left) == "dirichlet" && bc_to_string(q.bc.top) == "dirichlet");
        // ic.var is parsed but never applied (io.cpp:305 vs 347-360)
        SimConfig v = merged_config(std::nullopt, {"--ic.var=temperature", "--ic.path=a.nc"});
        CHECK(v.ic.var.empty() && v.ic.path == "a.nc");
    }
    std::remove(dev.c_str());
}

// ---- minimal CDF-5 reader for the round trip -----------------------------------------------------
struct Cdf5 {
    std::vector<unsigned char> b;
    size_t p = 0;
    int64_t i64() {
        int64_t v = 0;
        for (int k = 0; k < 8; ++k) v = (v << 8) | b[p++];
        return v;
    }
    uint32_t u32() {
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) v = (v << 8) | b[p++];
        return v;
    }
    std::string name() {
        const int64_t n = i64();
        std::string s(b.begin() + static_cast<long>(p), b.begin() + static_cast<long>(p) + n);
        p += static_cast<size_t>((n + 3) / 4 * 4);
        return s;
    }
};

static void snapshot_round_trip() {
    auto make_decomp = [](int nxg, int nyg, int nxl, int nyl) {
        Decomp2D d{};
        d.nx_global = nxg;
        d.ny_global = nyg;
        d.nx_local = nxl;
        d.ny_local = nyl;
        return d;
    };
    const std::string fname = "field_b200.nc";
    SimConfig cfg;
    cfg.nx = 3;
    cfg.ny = 2;
    cfg.dt = 0.123;
    cfg.steps = 10;
    cfg.D = 0.01;
    cfg.vx = 1.0;
    cfg.vy = -1.0;
    cfg.bc.right = BCType::Neumann;
    cfg.bc.bottom = BCType::Periodic;
    auto dec = make_decomp(3, 2, 3, 2);
    Field f(3, 2, 1, 1.0, 1.0);
    f.fill(-7.0);  // ghosts must not leak into the file
    for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 3; ++i) f.at(i + 1, j + 1) = 1.5 + i + 10 * j;
    int ncid = 0, varid = 0;
    CHECK(open_netcdf_parallel(fname, dec, cfg, MPI_COMM_WORLD, ncid, varid) == 0);
    CHECK(write_field_netcdf(ncid, varid, f, dec, 0));
    f.at(1, 1) = 99.0;
    CHECK(write_field_netcdf(ncid, varid, f, dec, 1));
    close_netcdf_parallel(ncid);

    Cdf5 r;
    {
        std::ifstream in(fname, std::ios::binary);
        r.b.assign(std::istreambuf_iterator<char>(in), {});
    }
    CHECK(r.b.size() > 64 && r.b[0] == 'C' && r.b[1] == 'D' && r.b[2] == 'F' && r.b[3] == 5);
    r.p = 4;
    CHECK(r.i64() == 2);                      // numrecs
    CHECK(r.u32() == 0x0A && r.i64() == 3);   // dimensions
    CHECK(r.name() == "time" && r.i64() == 0);
    CHECK(r.name() == "y" && r.i64() == 2);
    CHECK(r.name() == "x" && r.i64() == 3);
    CHECK(r.u32() == 0x0C && r.i64() == 7);   // global attributes
    std::vector<std::pair<std::string, std::string>> att;
    for (int k = 0; k < 7; ++k) {
        std::string n = r.name();
        CHECK(r.u32() == 2);
        att.emplace_back(n, r.name());
    }
    CHECK(att[0].first == "description" && att[0].second == "climate-sim-mpi-cpp");
    CHECK(att[1].first == "grid" && att[1].second == "3 x 2");
    CHECK(att[2].first == "dt" && att[2].second == "0.123000");
    CHECK(att[3].second == "10" && att[4].second == "0.010000");
    CHECK(att[5].first == "velocity" && att[5].second == "(1.000000,-1.000000)");
    CHECK(att[6].second == "left=dirichlet right=neumann bottom=periodic top=dirichlet");
    CHECK(r.u32() == 0x0B && r.i64() == 1);   // variables
    CHECK(r.name() == "u" && r.i64() == 3 && r.i64() == 0 && r.i64() == 1 && r.i64() == 2);
    CHECK(r.u32() == 0 && r.i64() == 0);      // no variable attributes
    CHECK(r.u32() == 6);                      // NC_DOUBLE
    CHECK(r.i64() == 3 * 2 * 8);              // vsize
    const int64_t begin = r.i64();
    CHECK(begin % 4 == 0 && static_cast<size_t>(begin) + 2 * 48 <= r.b.size());
    auto dbl = [&](int rec, int k) {
        r.p = static_cast<size_t>(begin) + static_cast<size_t>(rec) * 48 + static_cast<size_t>(k) * 8;
        const int64_t bits = r.i64();
        double d;
        std::memcpy(&d, &bits, 8);
        return d;
    };
    CHECK(dbl(0, 0) == 1.5 && dbl(0, 2) == 3.5 && dbl(0, 3) == 11.5 && dbl(0, 5) == 13.5);
    CHECK(dbl(1, 0) == 99.0 && dbl(1, 5) == 13.5);
    std::remove(fname.c_str());
}

// `test_driver --write-file <path> <2|5> <nx> <ny> <nrec>`: a complete file on the CPU from the writer's own
// header builder — u[t][y][x] = 1e6 t + 1e3 y + x + 0.25, big-endian — for readers the tests did not write
// (scipy.io.netcdf_file on the CDF-2 flavour, tests/test_driver.py).
static int write_file(const std::string& path, int format, int nx, int ny, int nrec) {
    const std::vector<std::pair<std::string, std::string>> attrs = {
        {"description", "climate-sim-mpi-cpp"}, {"grid", std::to_string(nx) + " x " + std::to_string(ny)},
        {"dt", std::to_string(0.1)}, {"odd", "abcde"}};
    int64_t begin = 0;
    const auto head = netcdf_header_bytes(format, nx, ny, nrec, attrs, &begin);
    std::vector<unsigned char> file(static_cast<size_t>(begin), 0);
    std::memcpy(file.data(), head.data(), head.size());
    for (int t = 0; t < nrec; ++t)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const double v = 1e6 * t + 1e3 * y + x + 0.25;
                uint64_t bits;
                std::memcpy(&bits, &v, 8);
                for (int sh = 56; sh >= 0; sh -= 8) file.push_back(static_cast<unsigned char>(bits >> sh));
            }
    std::ofstream out(path, std::ios::binary);
    out.write(reinterpret_cast<const char*>(file.data()), static_cast<std::streamsize>(file.size()));
    return out.good() ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc == 7 && std::string(argv[1]) == "--write-file")
        return write_file(argv[2], std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), std::atoi(argv[6]));
    config_tests();
    if (argc > 1 && std::string(argv[1]) == "--gpu") {
        MPI_Init(&argc, &argv);
        snapshot_round_trip();
        MPI_Finalize();
    }
    std::printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "ALL PASS", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
