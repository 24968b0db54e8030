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

struct ICConfig {
    std::string mode = "preset";
    std::string preset = "gaussian_hotspot";
    double A = 1.0;
    double sigma_frac = 0.05;
    double xc_frac = 0.5;
    double yc_frac = 0.5;
    std::string path;
    std::string var;
};

struct SimConfig {
    int nx = 256, ny = 256;
    double dx = 1.0, dy = 1.0;

    double D = 0.0;
    double vx = 0.0, vy = 0.0;

    double dt = 0.1;
    int steps = 100;
    int out_every = 50;

    BCConfig bc;

    std::string output_prefix = "snap";

    ICConfig ic{};

    void validate() const;  // throws std::runtime_error with the reference's messages (io.cpp:58-69)
};

struct CLIOverrides {
    std::optional<int> nx, ny;
    std::optional<double> dx, dy;

    std::optional<double> D, vx, vy;

    std::optional<double> dt;
    std::optional<int> steps, out_every;

    std::optional<BCType> bc_left, bc_right, bc_bottom, bc_top;

    std::optional<std::string> output_prefix;

    struct {
        std::optional<std::string> mode, preset, path, format, var;
        std::optional<double> A, sigma_frac, xc_frac, yc_frac;
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
void write_metadata_netcdf(int ncid, const SimConfig& cfg);

void apply_initial_condition(const Decomp2D& dec, Field& u, const SimConfig& cfg);
