// oracle/stub/mini_mpi.cpp — thread-per-rank emulation of the MPI subset declared in stub/mpi.h.
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_port.c header).  Not a port of any MPI library:
// eager buffered sends into a mailbox keyed by (source, dest, tag); receives complete in Waitall.
#include "mpi.h"

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace {

struct DType {
    int count = 1, blocklen = 1, stride = 1;  // in doubles
};
struct Req {
    bool is_recv = false;
    double* buf = nullptr;
    DType dt;
    int count = 0, peer = MPI_PROC_NULL, tag = 0;
};

int g_world = 1;
int g_cart_dims[2] = {1, 1};
thread_local int t_rank = 0;
thread_local std::vector<DType> t_types;  // handle = index + 2 (0 invalid, 1 = MPI_DOUBLE)
thread_local std::vector<Req> t_reqs;
thread_local int t_live_types = 0;

std::mutex g_mu;
std::condition_variable g_cv;
std::map<std::tuple<int, int, int>, std::deque<std::vector<double>>> g_mail;

// sense-reversing barrier over g_world threads
std::mutex b_mu;
std::condition_variable b_cv;
int b_count = 0;
unsigned b_gen = 0;

DType lookup(MPI_Datatype t) {
    if (t == MPI_DOUBLE) return DType{};
    return t_types.at(static_cast<size_t>(t - 2));
}

void pack(const double* src, const DType& d, int count, std::vector<double>& out) {
    // extent of a vector type = ((count-1)*stride + blocklen) elements; `count` repeats of it
    const size_t extent = static_cast<size_t>(d.count - 1) * d.stride + d.blocklen;
    for (int c = 0; c < count; ++c)
        for (int k = 0; k < d.count; ++k)
            for (int b = 0; b < d.blocklen; ++b)
                out.push_back(src[c * extent + static_cast<size_t>(k) * d.stride + b]);
}
void unpack(double* dst, const DType& d, int count, const std::vector<double>& in) {
    const size_t extent = static_cast<size_t>(d.count - 1) * d.stride + d.blocklen;
    size_t p = 0;
    for (int c = 0; c < count; ++c)
        for (int k = 0; k < d.count; ++k)
            for (int b = 0; b < d.blocklen; ++b)
                dst[c * extent + static_cast<size_t>(k) * d.stride + b] = in[p++];
}

}  // namespace

extern "C" {

void mini_mpi_set_world(int size) {
    g_world = size;
    std::lock_guard<std::mutex> lk(g_mu);
    g_mail.clear();
}
void mini_mpi_set_rank(int rank) {
    t_rank = rank;
    t_types.clear();
    t_reqs.clear();
    t_live_types = 0;
}
void mini_mpi_barrier(void) {
    std::unique_lock<std::mutex> lk(b_mu);
    const unsigned gen = b_gen;
    if (++b_count == g_world) {
        b_count = 0;
        ++b_gen;
        b_cv.notify_all();
    } else {
        b_cv.wait(lk, [&] { return gen != b_gen; });
    }
}

int MPI_Comm_size(MPI_Comm, int* size) {
    *size = g_world;
    return MPI_SUCCESS;
}
int MPI_Comm_rank(MPI_Comm, int* rank) {
    *rank = t_rank;
    return MPI_SUCCESS;
}

// MPI-3.1 §7.5.2: dims as close to each other as possible, in non-increasing order.
int MPI_Dims_create(int nnodes, int ndims, int dims[]) {
    if (ndims != 2) return 1;
    int b = 1;
    for (int d = 1; static_cast<long>(d) * d <= nnodes; ++d)
        if (nnodes % d == 0) b = d;
    dims[0] = nnodes / b;
    dims[1] = b;
    return MPI_SUCCESS;
}
// Row-major rank order, reorder ignored (the reference passes reorder=0, periods={0,0}).
int MPI_Cart_create(MPI_Comm, int ndims, const int dims[], const int*, int, MPI_Comm* comm_cart) {
    if (ndims != 2) return 1;
    g_cart_dims[0] = dims[0];  // every rank writes the same values
    g_cart_dims[1] = dims[1];
    *comm_cart = 2;
    return MPI_SUCCESS;
}
int MPI_Cart_coords(MPI_Comm, int rank, int, int coords[]) {
    coords[0] = rank / g_cart_dims[1];
    coords[1] = rank % g_cart_dims[1];
    return MPI_SUCCESS;
}
int MPI_Cart_shift(MPI_Comm, int direction, int disp, int* rank_source, int* rank_dest) {
    int c[2] = {t_rank / g_cart_dims[1], t_rank % g_cart_dims[1]};
    auto at = [&](int v) -> int {
        if (v < 0 || v >= g_cart_dims[direction]) return MPI_PROC_NULL;
        int cc[2] = {c[0], c[1]};
        cc[direction] = v;
        return cc[0] * g_cart_dims[1] + cc[1];
    };
    *rank_source = at(c[direction] - disp);
    *rank_dest = at(c[direction] + disp);
    return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* comm) {
    *comm = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype, MPI_Datatype* newtype) {
    t_types.push_back(DType{count, blocklength, stride});
    ++t_live_types;
    *newtype = static_cast<int>(t_types.size()) + 1;
    return MPI_SUCCESS;
}
int MPI_Type_contiguous(int count, MPI_Datatype, MPI_Datatype* newtype) {
    t_types.push_back(DType{1, count, count});
    ++t_live_types;
    *newtype = static_cast<int>(t_types.size()) + 1;
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype* t) {
    *t = 0;
    if (--t_live_types <= 0) {
        t_live_types = 0;
        t_types.clear();
    }
    return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm,
              MPI_Request* request) {
    Req r;
    r.is_recv = true;
    r.buf = static_cast<double*>(buf);
    r.dt = lookup(datatype);
    r.count = count;
    r.peer = source;
    r.tag = tag;
    t_reqs.push_back(r);
    *request = static_cast<int>(t_reqs.size()) - 1;
    return MPI_SUCCESS;
}
int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm,
              MPI_Request* request) {
    std::vector<double> payload;
    pack(static_cast<const double*>(buf), lookup(datatype), count, payload);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_mail[{t_rank, dest, tag}].push_back(std::move(payload));
    }
    g_cv.notify_all();
    Req r;
    t_reqs.push_back(r);
    *request = static_cast<int>(t_reqs.size()) - 1;
    return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status*) {
    for (int k = 0; k < count; ++k) {
        Req& r = t_reqs.at(static_cast<size_t>(reqs[k]));
        if (!r.is_recv) continue;
        std::vector<double> payload;
        {
            std::unique_lock<std::mutex> lk(g_mu);
            auto key = std::make_tuple(r.peer, t_rank, r.tag);
            g_cv.wait(lk, [&] {
                auto it = g_mail.find(key);
                return it != g_mail.end() && !it->second.empty();
            });
            auto& q = g_mail[key];
            payload = std::move(q.front());
            q.pop_front();
        }
        unpack(r.buf, r.dt, r.count, payload);
    }
    t_reqs.clear();
    return MPI_SUCCESS;
}

}  // extern "C"
