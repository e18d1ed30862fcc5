// outfile.cu — the output / checkpoint file of the Λ-iteration, written as HDF5 without the HDF5 library.
//
// Replaces create_output_file (reference src/io.jl:159-190 regular grid, :196-225 Voronoi sites) and the write_to_file
// methods (:57-157): one flat root group holding contiguous little-endian datasets with the reference's names and shapes, so
// that the reference's own readers (src/recover_simulation.jl:213-277, python/*.py through h5py) keep working when the Julia
// host hands the I/O to the library.  The image has no libhdf5, so the file is produced byte by byte from the HDF5 File Format
// Specification, in the oldest and most widely readable variant (what libhdf5 1.6/1.8 writes by default): version 0
// superblock, version 1 object headers, a symbol-table group (local heap + one v1 B-tree node + one symbol-table node with
// the file's leaf K raised to 16 so that all entries fit one node).  tests/h5mini_reader.py — an independent reader pinned
// against a file written by the real library — checks every structure.
//
// Dimension order: HDF5.jl stores a Julia (column-major) array with its dimensions reversed, so the Julia matrix
// source_function (nλ, n_sites) is the HDF5 dataset (n_sites, nλ) over the same bytes; the tables below list Julia shapes.
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <string>
#include <vector>
#include "vrt_internal.h"

namespace vrt {

constexpr uint64_t H5_UNDEF = 0xFFFFFFFFFFFFFFFFull;
constexpr int H5_LEAF_K = 16, H5_INTERNAL_K = 16;

struct OutDataset {
    std::string name;
    std::vector<uint64_t> dims;   // HDF5 order (slowest first) = reversed Julia shape
    bool is_int = false;
    uint64_t header = 0, data = 0, nbytes = 0, heap_off = 0;
};

struct ByteBuf {
    std::vector<unsigned char> b;
    void u8(unsigned v) { b.push_back((unsigned char)v); }
    void u16(unsigned v) { u8(v & 0xff); u8((v >> 8) & 0xff); }
    void u32(uint32_t v) { for (int i = 0; i < 4; i++) u8((v >> (8 * i)) & 0xff); }
    void u64(uint64_t v) { for (int i = 0; i < 8; i++) u8((unsigned)((v >> (8 * i)) & 0xff)); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void bytes(const void* p, size_t n) { b.insert(b.end(), (const unsigned char*)p, (const unsigned char*)p + n); }
    void pad_to(size_t n) { if (b.size() < n) zeros(n - b.size()); }
    size_t size() const { return b.size(); }
};

static uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

}  // namespace vrt

struct vrt_outfile {
    int fd = -1;
    std::string path;
    std::vector<vrt::OutDataset> ds;
    int64_t n_sites = 0, nlam = 0;
    const vrt::OutDataset* find(const char* name) const {
        for (auto& d : ds)
            if (d.name == name) return &d;
        return nullptr;
    }
    ~vrt_outfile() {
        if (fd >= 0) close(fd);
    }
};

namespace vrt {

static int io_fail(const char* what, const std::string& path) {
    set_error("%s '%s': %s", what, path.c_str(), strerror(errno));
    return VRT_E_IO;
}

static int pwrite_all(int fd, const void* p, size_t n, uint64_t off, const std::string& path) {
    const unsigned char* q = (const unsigned char*)p;
    while (n) {
        ssize_t w = pwrite(fd, q, std::min<size_t>(n, (size_t)1 << 30), (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            return io_fail("write to", path);
        }
        q += w; off += (uint64_t)w; n -= (size_t)w;
    }
    return VRT_OK;
}

// object header (version 1) of one dataset
static void dataset_header(ByteBuf& o, const OutDataset& d) {
    ByteBuf m;
    auto msg = [&](unsigned type, unsigned flags, const ByteBuf& body) {
        const size_t padded = (size_t)align_up(body.size(), 8);
        m.u16(type); m.u16((unsigned)padded); m.u8(flags); m.zeros(3);
        m.bytes(body.b.data(), body.size());
        m.zeros(padded - body.size());
    };
    {   // fill value (version 1): allocation late, written if set, "defined" with size 0 = the type's default — what libhdf5 writes
        ByteBuf b; b.u8(1); b.u8(2); b.u8(2); b.u8(1); b.u32(0);
        msg(0x0005, 1, b);
    }
    {   // datatype
        ByteBuf b;
        if (d.is_int) {       // class 0 fixed-point, version 1; little-endian, two's complement signed; 64 bits at offset 0
            b.u8(0x10); b.u8(0x08); b.u8(0); b.u8(0); b.u32(8); b.u16(0); b.u16(64);
        } else {              // class 1 floating-point, version 1; IEEE binary64 little-endian
            b.u8(0x11); b.u8(0x20); b.u8(0x3f); b.u8(0); b.u32(8);
            b.u16(0); b.u16(64); b.u8(52); b.u8(11); b.u8(0); b.u8(52); b.u32(1023);
        }
        msg(0x0003, 1, b);
    }
    {   // dataspace (version 1, no maximum dimensions)
        ByteBuf b; b.u8(1); b.u8((unsigned)d.dims.size()); b.u8(0); b.zeros(5);
        for (uint64_t v : d.dims) b.u64(v);
        msg(0x0001, 0, b);
    }
    {   // data layout (version 3, contiguous)
        ByteBuf b; b.u8(3); b.u8(1); b.u64(d.data); b.u64(d.nbytes);
        msg(0x0008, 1, b);
    }
    o.u8(1); o.u8(0); o.u16(4); o.u32(1); o.u32((uint32_t)m.size()); o.zeros(4);
    o.bytes(m.b.data(), m.size());
}

static size_t dataset_header_size(const OutDataset& d) {
    ByteBuf o;
    dataset_header(o, d);
    return o.size();
}

// lays the file out and writes all metadata; raw data areas are left as holes (the reference creates them `undef`)
static int outfile_create(const char* path, std::vector<OutDataset> ds, const char* zero_name, vrt_outfile** out) {
    if (!path || !out) return VRT_E_INVALID;
    *out = nullptr;
    if ((int)ds.size() > 2 * H5_LEAF_K) {
        set_error("output file: too many datasets for one symbol-table node");
        return VRT_E_INVALID;
    }
    std::sort(ds.begin(), ds.end(), [](const OutDataset& a, const OutDataset& b) { return strcmp(a.name.c_str(), b.name.c_str()) < 0; });
    // local heap data segment: the empty name at offset 0, then the link names, 8-byte aligned, then one free block
    uint64_t hoff = 8;
    for (auto& d : ds) {
        d.heap_off = hoff;
        hoff += align_up(d.name.size() + 1, 8);
    }
    const uint64_t heap_used = hoff, heap_size = align_up(heap_used + 16, 64);
    const uint64_t a_root = 96, a_heap = a_root + 16 + 32, a_heapdata = a_heap + 32, a_btree = a_heapdata + heap_size;
    const uint64_t btree_size = 24 + (uint64_t)(2 * H5_INTERNAL_K) * 16 + 8, a_snod = a_btree + btree_size;
    const uint64_t snod_size = 8 + (uint64_t)(2 * H5_LEAF_K) * 40;
    uint64_t pos = a_snod + snod_size;
    for (auto& d : ds) {
        d.nbytes = 8;
        for (uint64_t v : d.dims) d.nbytes *= v;
        d.header = pos;
        pos += dataset_header_size(d);
    }
    for (auto& d : ds) {   // raw data: page-aligned so that the big arrays stream with aligned writes
        pos = align_up(pos, d.nbytes >= (1u << 20) ? 4096 : 8);
        d.data = pos;
        pos += d.nbytes;
    }
    const uint64_t eof = pos;

    ByteBuf o;
    // ---- superblock, version 0
    o.bytes("\x89HDF\r\n\x1a\n", 8);
    o.u8(0); o.u8(0); o.u8(0); o.u8(0); o.u8(0);   // superblock, free-space, root symbol table entry versions; reserved; shared header version
    o.u8(8); o.u8(8); o.u8(0);                     // size of offsets, size of lengths, reserved
    o.u16(H5_LEAF_K); o.u16(H5_INTERNAL_K);
    o.u32(0);                                      // file consistency flags
    o.u64(0); o.u64(H5_UNDEF); o.u64(eof); o.u64(H5_UNDEF);   // base, free-space info, end of file, driver information
    o.u64(0); o.u64(a_root); o.u32(1); o.u32(0); o.u64(a_btree); o.u64(a_heap);   // root symbol table entry (cached B-tree / heap)
    o.pad_to(a_root);
    // ---- root group object header: symbol table message + NIL
    o.u8(1); o.u8(0); o.u16(2); o.u32(1); o.u32(32); o.zeros(4);
    o.u16(0x0011); o.u16(16); o.u8(1); o.zeros(3); o.u64(a_btree); o.u64(a_heap);
    o.u16(0x0000); o.u16(0); o.u8(0); o.zeros(3);
    // ---- local heap
    o.pad_to(a_heap);
    o.bytes("HEAP", 4); o.u8(0); o.zeros(3); o.u64(heap_size); o.u64(heap_used); o.u64(a_heapdata);
    o.pad_to(a_heapdata);
    o.zeros(8);
    for (auto& d : ds) {
        o.bytes(d.name.c_str(), d.name.size() + 1);
        o.pad_to(a_heapdata + d.heap_off + align_up(d.name.size() + 1, 8));
    }
    o.u64(1); o.u64(heap_size - heap_used);        // the free block: no next block (1 = H5HL_FREE_NULL), its size
    o.pad_to(a_btree);
    // ---- B-tree node (group node, leaf level): one child = the symbol table node
    o.bytes("TREE", 4); o.u8(0); o.u8(0); o.u16(1); o.u64(H5_UNDEF); o.u64(H5_UNDEF);
    o.u64(0); o.u64(a_snod); o.u64(ds.empty() ? 0 : ds.back().heap_off);
    o.pad_to(a_snod);
    // ---- symbol table node
    o.bytes("SNOD", 4); o.u8(1); o.u8(0); o.u16((unsigned)ds.size());
    for (auto& d : ds) {
        o.u64(d.heap_off); o.u64(d.header); o.u32(0); o.u32(0); o.zeros(16);
    }
    o.pad_to(a_snod + snod_size);
    for (auto& d : ds) {
        if (o.size() != d.header) {
            set_error("output file: internal layout error");
            return VRT_E_STATE;
        }
        dataset_header(o, d);
    }

    vrt_outfile* f = new vrt_outfile();
    struct Guard { vrt_outfile* f; ~Guard() { delete f; } } guard{f};
    f->path = path;
    f->fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (f->fd < 0) return io_fail("cannot create", f->path);
    if (ftruncate(f->fd, (off_t)eof) != 0) return io_fail("cannot size", f->path);
    VRT_TRY(pwrite_all(f->fd, o.b.data(), o.size(), 0, f->path));
    f->ds = ds;
    (void)zero_name;   // `convergence` is created as zeros (io.jl:216): the hole ftruncate leaves reads as zeros already
    guard.f = nullptr;
    *out = f;
    return VRT_OK;
}

static OutDataset dset(const char* name, std::initializer_list<int64_t> julia_shape, bool is_int = false) {
    OutDataset d;
    d.name = name;
    d.is_int = is_int;
    for (auto it = julia_shape.end(); it != julia_shape.begin();) d.dims.push_back((uint64_t)*--it);
    return d;
}

// part of a dataset, from host memory (solver.cu streams the device state through a pinned staging buffer)
int outfile_write_at(vrt_outfile* f, const char* name, uint64_t offset, const void* host, size_t nbytes) {
    const OutDataset* d = f ? f->find(name) : nullptr;
    if (!d || offset + nbytes > d->nbytes) {
        set_error("output file: bad partial write to '%s'", name ? name : "?");
        return VRT_E_INVALID;
    }
    return pwrite_all(f->fd, host, nbytes, d->data + offset, f->path);
}
int outfile_shape(const vrt_outfile* f, int64_t* n_sites, int64_t* nlam) {
    if (!f) return VRT_E_INVALID;
    *n_sites = f->n_sites;
    *nlam = f->nlam;
    return VRT_OK;
}

}  // namespace vrt

using namespace vrt;

extern "C" {

int vrt_output_create(const char* path, int64_t nlam, int64_t n_sites, int64_t maxiter, vrt_outfile** out) {
    if (nlam <= 0 || n_sites <= 0 || maxiter < 0) {
        set_error("vrt_output_create: bad sizes");
        return VRT_E_INVALID;
    }
    // create_output_file(output_path, nλ, n_sites, maxiter), io.jl:196-225
    std::vector<OutDataset> ds = {
        dset("source_function", {nlam, n_sites}), dset("populations", {n_sites, 3}), dset("positions", {3, n_sites}),
        dset("temperature", {n_sites}), dset("hydrogen_populations", {n_sites}), dset("electron_density", {n_sites}),
        dset("velocity_z", {n_sites}), dset("velocity_x", {n_sites}), dset("velocity_y", {n_sites}), dset("boundaries", {6}),
        dset("convergence", {maxiter + 1}), dset("n_bb", {1}, true), dset("n_bf", {1}, true), dset("wavelength", {nlam}),
        dset("line_center", {1}), dset("time", {1})};
    VRT_TRY(outfile_create(path, ds, "convergence", out));
    (*out)->n_sites = n_sites;
    (*out)->nlam = nlam;
    return VRT_OK;
}

int vrt_output_create_regular(const char* path, int64_t nlam, int64_t nz, int64_t nx, int64_t ny, int64_t maxiter, vrt_outfile** out) {
    if (nlam <= 0 || nz <= 0 || nx <= 0 || ny <= 0 || maxiter < 0) {
        set_error("vrt_output_create_regular: bad sizes");
        return VRT_E_INVALID;
    }
    // create_output_file(output_path, nλ, (nz, nx, ny), maxiter), io.jl:159-190
    std::vector<OutDataset> ds = {
        dset("source_function", {nlam, nz, nx, ny}), dset("populations", {nz, nx, ny, 3}), dset("z", {nz}), dset("x", {nx}), dset("y", {ny}),
        dset("temperature", {nz, nx, ny}), dset("hydrogen_populations", {nz, nx, ny}), dset("electron_density", {nz, nx, ny}),
        dset("velocity_z", {nz, nx, ny}), dset("velocity_x", {nz, nx, ny}), dset("velocity_y", {nz, nx, ny}),
        dset("convergence", {maxiter + 1}), dset("n_bb", {1}, true), dset("n_bf", {1}, true), dset("wavelength", {nlam}),
        dset("line_center", {1}), dset("time", {1})};
    VRT_TRY(outfile_create(path, ds, "convergence", out));
    (*out)->n_sites = nz * nx * ny;
    (*out)->nlam = nlam;
    return VRT_OK;
}

int vrt_output_dataset_size(const vrt_outfile* f, const char* name, int64_t* nbytes) {
    if (!f || !name || !nbytes) return VRT_E_INVALID;
    const OutDataset* d = f->find(name);
    if (!d) {
        set_error("output file: no dataset '%s'", name);
        return VRT_E_INVALID;
    }
    *nbytes = (int64_t)d->nbytes;
    return VRT_OK;
}

// write_to_file(array, output_path): the whole dataset `name`; data may be a host or a device pointer
int vrt_output_write(vrt_outfile* f, const char* name, const void* data, int64_t nbytes) {
    if (!f || !name || !data) return VRT_E_INVALID;
    const OutDataset* d = f->find(name);
    if (!d) {
        set_error("output file: no dataset '%s'", name);
        return VRT_E_INVALID;
    }
    if ((uint64_t)nbytes != d->nbytes) {
        set_error("output file: dataset '%s' holds %llu bytes, %lld given", name, (unsigned long long)d->nbytes, (long long)nbytes);
        return VRT_E_INVALID;
    }
    if (!is_device_ptr(data)) return pwrite_all(f->fd, data, (size_t)nbytes, d->data, f->path);
    const size_t chunk = (size_t)64 << 20;
    std::vector<unsigned char> host(std::min<size_t>(chunk, (size_t)nbytes));
    for (size_t o = 0; o < (size_t)nbytes; o += chunk) {
        const size_t m = std::min(chunk, (size_t)nbytes - o);
        VRT_CUDA(cudaMemcpy(host.data(), (const unsigned char*)data + o, m, cudaMemcpyDeviceToHost));
        VRT_TRY(pwrite_all(f->fd, host.data(), m, d->data + o, f->path));
    }
    return VRT_OK;
}

// write_to_file(difference, iteration, output_path), io.jl:129-135: convergence[iteration] = difference (1-based)
int vrt_output_write_convergence(vrt_outfile* f, int64_t iteration, double difference) {
    if (!f) return VRT_E_INVALID;
    const OutDataset* d = f->find("convergence");
    if (!d || iteration < 1 || (uint64_t)iteration * 8 > d->nbytes) {
        set_error("output file: convergence index %lld out of range", (long long)iteration);
        return VRT_E_INVALID;
    }
    return pwrite_all(f->fd, &difference, 8, d->data + (uint64_t)(iteration - 1) * 8, f->path);
}

int vrt_output_close(vrt_outfile* f) {
    if (!f) return VRT_E_INVALID;
    int rc = VRT_OK;
    if (f->fd >= 0 && fsync(f->fd) != 0) rc = io_fail("cannot flush", f->path);
    delete f;
    return rc;
}

}  // extern "C"
