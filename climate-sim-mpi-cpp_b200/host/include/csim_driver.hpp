// csim_driver.hpp — the callers and data formats either side of the timestep path (SURVEY.md §8f,
// rows N1–N4), with the reference's names so that src/main.cpp and tests/simulation/unit/test_io.cpp
// compile against this build:
//   ICConfig, SimConfig, CLIOverrides, load_yaml_file, parse_cli_overrides, merged_config,
//   bc_from_string, bc_to_string                         include/io.hpp:10-68, src/io.cpp:30-376
//   open_netcdf_parallel, write_field_netcdf, close_netcdf_parallel, write_metadata_netcdf
//                                                         include/io.hpp:70-81, src/io.cpp:378-448
//   apply_initial_condition                               include/init.hpp:6, src/init.cpp:12-47
// yaml-cpp and PnetCDF are not available here, so the YAML subset the reference's files use is
// parsed by a small reader (config.cpp) and snapshots are written by a native CDF-5 writer
// (cdf5.cpp): the same on-disk format PnetCDF produces with NC_64BIT_DATA — one record variable
// u(time, y, x) of big-endian doubles plus seven global text attributes.
#pragma once
#include <optional>
#include <string>
#include <vector>

#include "csim_dropin.hpp"

// Initial-condition block of the configuration (include/io.hpp:10-19).
struct ICConfig {
    std::string mode = "preset";              // "preset" | "file" ("file" throws, init.cpp:44-46)
    std::string preset = "gaussian_hotspot";  // or "constant_zero"
    double A = 1.0;                           // amplitude
    double sigma_frac = 0.05;                 // sigma = sigma_frac * min(Lx, Ly)
    double xc_frac = 0.5;                     // centre, as a fraction of Lx
    double yc_frac = 0.5;                     // centre, as a fraction of Ly
    std::string path;                         // parsed, unused upstream
    std::string var;                          // parsed from YAML only; the CLI value is dropped upstream
};

// Everything the run needs (include/io.hpp:21-39); defaults are the reference's.
struct SimConfig {
    int nx = 256;  // global grid
    int ny = 256;
    double dx = 1.0;  // spacing
    double dy = 1.0;
    double D = 0.0;  // diffusivity
    double vx = 0.0;  // constant velocity
    double vy = 0.0;
    double dt = 0.1;  // clamped to safe_dt by the driver
    int steps = 100;
    int out_every = 50;  // a frame at the start of every out_every-th step
    BCConfig bc;         // four sides, Dirichlet by default
    std::string output_prefix = "snap";  // parsed; the output path is fixed (main.cpp:87)
    ICConfig ic{};

    // throws std::runtime_error with the reference's messages (io.cpp:58-69)
    void validate() const;
};

// Values given on the command line (include/io.hpp:41-62); an empty optional keeps the YAML/default.
struct CLIOverrides {
    std::optional<int> nx;
    std::optional<int> ny;
    std::optional<double> dx;
    std::optional<double> dy;
    std::optional<double> D;
    std::optional<double> vx;
    std::optional<double> vy;
    std::optional<double> dt;
    std::optional<int> steps;
    std::optional<int> out_every;
    std::optional<BCType> bc_left;
    std::optional<BCType> bc_right;
    std::optional<BCType> bc_bottom;
    std::optional<BCType> bc_top;
    std::optional<std::string> output_prefix;
    struct IC {
        std::optional<std::string> mode;
        std::optional<std::string> preset;
        std::optional<std::string> path;
        std::optional<std::string> format;  // never parsed upstream either
        std::optional<std::string> var;
        std::optional<double> A;
        std::optional<double> sigma_frac;
        std::optional<double> xc_frac;
        std::optional<double> yc_frac;
    } ic;
};

SimConfig load_yaml_file(const std::string& path);
CLIOverrides parse_cli_overrides(const std::vector<std::string>& args);
SimConfig merged_config(const std::optional<std::string>& yaml_path, const std::vector<std::string>& cli_args);

BCType bc_from_string(const std::string& s);
std::string bc_to_string(BCType bc);

// Snapshot output.  `ncid` is a handle of this library's CDF-5 writer, not a PnetCDF id.
int open_netcdf_parallel(const std::string& filename, const Decomp2D& dec, const SimConfig& cfg, MPI_Comm comm,
                         int& ncid, int& varid);
bool write_field_netcdf(int ncid, int varid, const Field& f, const Decomp2D& dec, int step);
void close_netcdf_parallel(int ncid);
// Not in the reference: true once a frame of this file could not be written (the writer thread's
// failures are asynchronous; call after the time loop, before close_netcdf_parallel).  Drains the writer.
bool netcdf_write_failed(int ncid);
void write_metadata_netcdf(int ncid, const SimConfig& cfg);

// The file header the writer emits, as bytes (host only, no GPU): dims time(unlimited), y, x and one
// record variable u(time,y,x) of NC_DOUBLE with global text attributes.  format 5 = CDF-5 (what the
// reference's NC_CLOBBER|NC_64BIT_DATA creates, src/io.cpp:386); format 2 = CDF-2, the same grammar with
// 32-bit counts (CSIM_NETCDF_FORMAT=cdf2 makes open_netcdf_parallel write it), which readers without
// CDF-5 support — scipy.io.netcdf_file — open.  *data_begin receives the offset of record 0.
std::vector<unsigned char> netcdf_header_bytes(int format, int64_t nx_global, int64_t ny_global, int64_t numrecs,
                                               const std::vector<std::pair<std::string, std::string>>& attrs,
                                               int64_t* data_begin);

void apply_initial_condition(const Decomp2D& dec, Field& u, const SimConfig& cfg);
