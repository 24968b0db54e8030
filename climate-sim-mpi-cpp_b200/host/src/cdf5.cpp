// cdf5.cpp — snapshot output in NetCDF CDF-5 (64-bit data) without PnetCDF.
//
// Stands behind open_netcdf_parallel / write_field_netcdf / close_netcdf_parallel /
// write_metadata_netcdf (include/io.hpp:70-81, src/io.cpp:378-448).  The file is what the reference's
// ncmpi_create(NC_CLOBBER | NC_64BIT_DATA) + def_dim(time unlimited, y, x) + def_var(u, NC_DOUBLE) +
// seven global text attributes produce: magic "CDF\x05", 64-bit counts, big-endian doubles, one record
// per frame.  netCDF-C (integration_helpers.cpp:27-74) and netCDF4-python (visualization/io.py) read it.
//
// Data path per frame (the reference: host loop through Field::at + collective put, io.cpp:411-418):
//   GPU kernel de-halos and byte-swaps the tile → async D2H into one of two pinned buffers → a
//   writer thread waits for that copy only and pwrite()s the rows at this rank's {y_off, x_off}
//   window of the record.  The time loop keeps running behind it; the call blocks only when both
//   buffers are still in flight.  Ranks of one box write disjoint byte ranges of the same file;
//   rank 0 owns the header and patches numrecs on close.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include "csim_driver.hpp"

namespace {

// ---- big-endian header builder ---------------------------------------------------------------
struct Bytes {
    std::vector<unsigned char> b;
    bool wide = true;  // CDF-5: counts, lengths and sizes are 64-bit; CDF-1/2: 32-bit
    void count(int64_t v) {
        if (wide)
            i64(v);
        else  // a size that does not fit (vsize of a huge record) is written as 2^32-1, as the format says
            u32(v > 0xFFFFFFFFll - 3 ? 0xFFFFFFFFu : static_cast<uint32_t>(v));
    }
    void u32(uint32_t v) {
        for (int s = 24; s >= 0; s -= 8) b.push_back(static_cast<unsigned char>(v >> s));
    }
    void i64(int64_t v) {
        for (int s = 56; s >= 0; s -= 8) b.push_back(static_cast<unsigned char>(static_cast<uint64_t>(v) >> s));
    }
    void name(const std::string& s) {  // nelems + bytes padded to a multiple of 4
        count(static_cast<int64_t>(s.size()));
        b.insert(b.end(), s.begin(), s.end());
        while (b.size() % 4) b.push_back(0);
    }
};
constexpr uint32_t NC_DIMENSION = 0x0A, NC_VARIABLE = 0x0B, NC_ATTRIBUTE = 0x0C;
constexpr uint32_t NC_CHAR = 2, NC_DOUBLE = 6;
constexpr int64_t kDataAlign = 4096;

struct Job {
    int buf;
    int step;
    void* event;
};

struct Cdf5File {
    int fd = -1;
    bool owner = false;  // rank 0: wrote the header, patches numrecs
    int nx_global = 0, ny_global = 0;
    int64_t begin = 0, recsize = 0;
    int64_t numrecs = 0;
    int format = 5;  // 5: CDF-5 (the reference's NC_64BIT_DATA), 2: CDF-2 (CSIM_NETCDF_FORMAT=cdf2)
    std::vector<std::pair<std::string, std::string>> attrs;
    bool header_written = false;
    std::string path;
    // this rank's window
    int nx = 0, ny = 0, x_off = 0, y_off = 0;
    // async pipeline
    void* pinned[2] = {nullptr, nullptr};
    bool busy[2] = {false, false};
    std::deque<Job> jobs;
    std::mutex mu;
    std::condition_variable cv;
    std::thread writer;
    bool stop = false, failed = false;
    std::string error;
};

std::mutex g_mu;
std::vector<std::unique_ptr<Cdf5File>> g_files;  // ncid = index + 1

Cdf5File* lookup(int ncid) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (ncid < 1 || ncid > static_cast<int>(g_files.size()) || !g_files[static_cast<size_t>(ncid - 1)])
        return nullptr;
    return g_files[static_cast<size_t>(ncid - 1)].get();
}

// The header grammar of the classic NetCDF formats (same for CDF-1, 2 and 5 up to field widths):
//   magic numrecs dim_list gatt_list var_list
//   dim_list = NC_DIMENSION nelems [name dim_length]…      (ABSENT = ZERO ZERO)
//   att      = name nc_type nelems [values, padded to 4]
//   var      = name ndims [dimid]… vatt_list nc_type vsize begin
// CDF-5: every count/length/size/dimid is 64-bit; CDF-2: 32-bit, except `begin`, which is 64-bit in both.
std::vector<unsigned char> header_bytes(int format, int64_t nx_global, int64_t ny_global, int64_t numrecs,
                                        const std::vector<std::pair<std::string, std::string>>& attrs,
                                        int64_t* begin_out) {
    const int64_t recsize = nx_global * ny_global * 8;
    auto emit = [&](int64_t begin) {
        Bytes h;
        h.wide = format == 5;
        h.b = {'C', 'D', 'F', static_cast<unsigned char>(format)};
        h.count(numrecs);
        h.u32(NC_DIMENSION);
        h.count(3);
        h.name("time");
        h.count(0);  // record dimension
        h.name("y");
        h.count(ny_global);
        h.name("x");
        h.count(nx_global);
        if (attrs.empty()) {
            h.u32(0);
            h.count(0);
        } else {
            h.u32(NC_ATTRIBUTE);
            h.count(static_cast<int64_t>(attrs.size()));
            for (auto& a : attrs) {
                h.name(a.first);
                h.u32(NC_CHAR);
                h.count(static_cast<int64_t>(a.second.size()));
                h.b.insert(h.b.end(), a.second.begin(), a.second.end());
                while (h.b.size() % 4) h.b.push_back(0);
            }
        }
        h.u32(NC_VARIABLE);
        h.count(1);
        h.name("u");
        h.count(3);
        h.count(0);
        h.count(1);
        h.count(2);
        h.u32(0);  // no variable attributes
        h.count(0);
        h.u32(NC_DOUBLE);
        h.count(recsize);  // vsize: one record of u
        h.i64(begin);      // OFFSET: 64-bit in CDF-2 and CDF-5
        return h.b;
    };
    const size_t len = emit(0).size();
    const int64_t begin = (static_cast<int64_t>(len) + kDataAlign - 1) / kDataAlign * kDataAlign;
    if (begin_out) *begin_out = begin;
    return emit(begin);
}

std::vector<unsigned char> build_header(const Cdf5File& f, int64_t numrecs, int64_t* begin_out) {
    return header_bytes(f.format, f.nx_global, f.ny_global, numrecs, f.attrs, begin_out);
}

bool pwrite_all(int fd, const void* p, size_t n, int64_t off) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        const ssize_t w = ::pwrite(fd, c, n, static_cast<off_t>(off));
        if (w < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        c += w;
        n -= static_cast<size_t>(w);
        off += w;
    }
    return true;
}

void write_header(Cdf5File& f) {
    int64_t begin = 0;
    const auto h = build_header(f, f.numrecs, &begin);
    f.begin = begin;
    if (!pwrite_all(f.fd, h.data(), h.size(), 0)) throw std::runtime_error("cdf5: header write failed: " + f.path);
    f.header_written = true;
}

void writer_loop(Cdf5File* f) {
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(f->mu);
            f->cv.wait(lk, [&] { return f->stop || !f->jobs.empty(); });
            if (f->jobs.empty()) return;
            j = f->jobs.front();
            f->jobs.pop_front();
        }
        std::string why;
        bool ok = csim_event_wait(csim_host::default_context(), j.event) == CSIM_OK;
        if (!ok) why = csim_last_error();  // thread-local: only this thread can read it
        if (ok) {
            const char* src = static_cast<const char*>(f->pinned[j.buf]);
            const size_t row = static_cast<size_t>(f->nx) * 8;
            const int64_t rec0 = f->begin + static_cast<int64_t>(j.step) * f->recsize;
            if (f->nx == f->nx_global) {  // whole rows: one contiguous range
                ok = pwrite_all(f->fd, src, row * static_cast<size_t>(f->ny),
                                rec0 + static_cast<int64_t>(f->y_off) * f->nx_global * 8);
            } else {
                for (int r = 0; r < f->ny && ok; ++r)
                    ok = pwrite_all(f->fd, src + static_cast<size_t>(r) * row, row,
                                    rec0 + (static_cast<int64_t>(f->y_off + r) * f->nx_global + f->x_off) * 8);
            }
        }
        {
            std::lock_guard<std::mutex> lk(f->mu);
            f->busy[j.buf] = false;
            if (!ok && !f->failed) {
                f->failed = true;
                f->error = std::string("Rank write failed: ") + (why.empty() ? std::strerror(errno) : why.c_str());
            }
        }
        f->cv.notify_all();
    }
}

std::string to_s(double v) { return std::to_string(v); }  // "%f", as the reference's std::to_string

}  // namespace

std::vector<unsigned char> netcdf_header_bytes(int format, int64_t nx_global, int64_t ny_global, int64_t numrecs,
                                               const std::vector<std::pair<std::string, std::string>>& attrs,
                                               int64_t* data_begin) {
    if (format != 2 && format != 5) throw std::runtime_error("netcdf_header_bytes: format must be 2 or 5");
    return header_bytes(format, nx_global, ny_global, numrecs, attrs, data_begin);
}

void write_metadata_netcdf(int ncid, const SimConfig& cfg) {
    Cdf5File* f = lookup(ncid);
    if (!f) {
        std::cerr << "Error writing attribute description: bad file id\n";
        return;
    }
    if (f->header_written && !f->owner) return;
    f->attrs = {
        {"description", "climate-sim-mpi-cpp"},
        {"grid", std::to_string(cfg.nx) + " x " + std::to_string(cfg.ny)},
        {"dt", to_s(cfg.dt)},
        {"steps", std::to_string(cfg.steps)},
        {"D", to_s(cfg.D)},
        {"velocity", "(" + to_s(cfg.vx) + "," + to_s(cfg.vy) + ")"},
        {"boundary_conditions", "left=" + bc_to_string(cfg.bc.left) + " right=" + bc_to_string(cfg.bc.right) +
                                    " bottom=" + bc_to_string(cfg.bc.bottom) + " top=" + bc_to_string(cfg.bc.top)},
    };
}

int open_netcdf_parallel(const std::string& filename, const Decomp2D& dec, const SimConfig& cfg, MPI_Comm comm,
                         int& ncid, int& varid) {
    int rank = 0;
    MPI_Comm_rank(comm, &rank);
    auto f = std::make_unique<Cdf5File>();
    f->path = filename;
    f->owner = rank == 0;
    f->nx_global = dec.nx_global;
    f->ny_global = dec.ny_global;
    f->recsize = static_cast<int64_t>(dec.nx_global) * dec.ny_global * 8;
    if (const char* fmt = std::getenv("CSIM_NETCDF_FORMAT")) {
        if (std::strcmp(fmt, "cdf2") == 0)
            f->format = 2;
        else if (std::strcmp(fmt, "cdf5") != 0)
            throw std::runtime_error(std::string("CSIM_NETCDF_FORMAT must be cdf5 or cdf2, not ") + fmt);
    }
    f->nx = dec.nx_local;
    f->ny = dec.ny_local;
    f->x_off = dec.x_offset;
    f->y_off = dec.y_offset;
    Cdf5File* raw = f.get();
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_files.push_back(std::move(f));
        ncid = static_cast<int>(g_files.size());
    }
    varid = 0;
    write_metadata_netcdf(ncid, cfg);  // io.cpp:397, before enddef
    if (raw->owner) {
        raw->fd = ::open(filename.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);  // NC_CLOBBER
        if (raw->fd < 0) throw std::runtime_error("ncmpi_create: " + std::string(std::strerror(errno)) + ": " + filename);
        write_header(*raw);
    }
    MPI_Barrier(comm);  // the file exists before anyone else opens it
    if (!raw->owner) {
        raw->fd = ::open(filename.c_str(), O_RDWR);
        if (raw->fd < 0) throw std::runtime_error("ncmpi_create: " + std::string(std::strerror(errno)) + ": " + filename);
        int64_t begin = 0;
        build_header(*raw, 0, &begin);  // same attributes on every rank → same data offset
        raw->begin = begin;
        raw->header_written = true;
    }
    const size_t bytes = static_cast<size_t>(raw->nx) * static_cast<size_t>(raw->ny) * 8;
    for (void*& p : raw->pinned) csim_host::check(csim_host_alloc(bytes ? bytes : 8, &p));
    raw->writer = std::thread(writer_loop, raw);
    return 0;  // NC_NOERR
}

bool write_field_netcdf(int ncid, int /*varid*/, const Field& fld, const Decomp2D& dec, int step) {
    Cdf5File* f = lookup(ncid);
    if (!f || f->fd < 0) {
        std::cerr << "Rank write failed: bad file id\n";
        return false;
    }
    if (dec.nx_local != f->nx || dec.ny_local != f->ny || fld.nx_local != f->nx || fld.ny_local != f->ny) {
        std::cerr << "Rank write failed: tile does not match the file's decomposition\n";
        return false;
    }
    int buf = -1;
    {
        std::unique_lock<std::mutex> lk(f->mu);
        f->cv.wait(lk, [&] { return f->failed || !f->busy[0] || !f->busy[1]; });
        if (f->failed) {
            std::cerr << f->error << "\n";
            return false;
        }
        buf = f->busy[0] ? 1 : 0;
        f->busy[buf] = true;
    }
    // pack + byte swap on the compute stream, the PCIe copy on the copy stream: the steps queued after this
    // call run while the frame travels
    void* ev = nullptr;
    const int rc2 = csim_field_snapshot_async(fld.data.device_ro(), f->pinned[buf], 1, &ev);
    if (rc2 != CSIM_OK) {
        std::lock_guard<std::mutex> lk(f->mu);
        f->busy[buf] = false;
        std::cerr << "Rank write failed: " << csim_last_error() << "\n";
        return false;
    }
    {
        std::lock_guard<std::mutex> lk(f->mu);
        f->jobs.push_back(Job{buf, step, ev});
        if (step + 1 > f->numrecs) f->numrecs = step + 1;
    }
    f->cv.notify_all();
    return true;
}

bool netcdf_write_failed(int ncid) {
    Cdf5File* f = lookup(ncid);
    if (!f) return true;
    std::unique_lock<std::mutex> lk(f->mu);
    f->cv.wait(lk, [&] { return f->jobs.empty() && !f->busy[0] && !f->busy[1]; });
    if (f->failed) std::cerr << f->error << "\n";
    return f->failed;
}

void close_netcdf_parallel(int ncid) {
    Cdf5File* f = lookup(ncid);
    if (!f) return;
    {
        std::unique_lock<std::mutex> lk(f->mu);
        f->cv.wait(lk, [&] { return f->jobs.empty() && !f->busy[0] && !f->busy[1]; });
        f->stop = true;
    }
    f->cv.notify_all();
    if (f->writer.joinable()) f->writer.join();
    // every rank saw the same frame count; rank 0 records it (ncmpi_close flushes numrecs)
    const double n_mine = static_cast<double>(f->numrecs);
    double n = n_mine;  // separate send and receive buffers: aliasing them needs MPI_IN_PLACE with a real MPI
    MPI_Reduce(&n_mine, &n, 1, MPI_DOUBLE, MPI_MAX, 0, MPI_COMM_WORLD);
    if (f->owner && f->fd >= 0) {
        f->numrecs = static_cast<int64_t>(n);
        Bytes b;
        b.wide = f->format == 5;
        b.count(f->numrecs);
        pwrite_all(f->fd, b.b.data(), b.b.size(), 4);
    }
    if (f->fd >= 0) ::close(f->fd);
    f->fd = -1;
    for (void*& p : f->pinned) {
        if (p) csim_host_free(p);
        p = nullptr;
    }
    MPI_Barrier(MPI_COMM_WORLD);
    std::lock_guard<std::mutex> lk(g_mu);
    g_files[static_cast<size_t>(ncid - 1)].reset();
}

// ---- initial condition: src/init.cpp:35-47 ------------------------------------------------------
void apply_initial_condition(const Decomp2D& dec, Field& u, const SimConfig& cfg) {
    if (cfg.ic.mode != "preset") throw std::runtime_error("IC mode 'file' not supported in PnetCDF build.");
    int preset;
    if (cfg.ic.preset == "gaussian_hotspot")
        preset = 0;
    else if (cfg.ic.preset == "constant_zero")
        preset = 1;
    else
        throw std::runtime_error("Unknown IC preset: " + cfg.ic.preset);
    if (preset == 1) return;  // u already zero
    csim_decomp c;
    std::memset(&c, 0, sizeof c);
    c.nx_global = cfg.nx;  // ic_gaussian takes Lx, Ly from cfg (init.cpp:17-18), offsets from dec
    c.ny_global = cfg.ny;
    c.nx_local = u.nx_local;
    c.ny_local = u.ny_local;
    c.x_offset = dec.x_offset;
    c.y_offset = dec.y_offset;
    // On the device when the host libm's exp() is one the library reproduces bit for bit (no host loop, no
    // upload); CSIM_IC=host forces the host path.  Needs halo 1 tiles like everything on the fused path.
    const char* where = std::getenv("CSIM_IC");
    if (!(where && std::strcmp(where, "host") == 0) && csim_exp_variant() >= 0 && u.dx == cfg.dx && u.dy == cfg.dy) {
        csim_host::check(csim_initial_condition_device(u.data.device_rw(), &c, cfg.nx, cfg.ny, preset, cfg.ic.A,
                                                       cfg.ic.sigma_frac, cfg.ic.xc_frac, cfg.ic.yc_frac));
        return;
    }
    csim_host::check(csim_initial_condition_host(u.data.data(), &c, u.halo, cfg.nx, cfg.ny, cfg.dx, cfg.dy, preset,
                                                 cfg.ic.A, cfg.ic.sigma_frac, cfg.ic.xc_frac, cfg.ic.yc_frac));
}
