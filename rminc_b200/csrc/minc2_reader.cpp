// minc2_reader.cpp -- MINC 2 volumes read WITHOUT libminc / libhdf5 (the step in front of
// the hot path).
//
// Reference: every volume of a mincLm call is opened with miopen_volume and read slab by slab with
// miget_real_value_hyperslab(..., MI_TYPE_DOUBLE, ...) (src/modelling_functions.c:806-829, 1037-1055;
// src/minc_reader.c:294-331; R/minc_interface.R:1031-1077): libminc asks HDF5 for the stored voxels (usually 16-bit,
// chunked and deflated), converts them to doubles with the slice's real range on one CPU thread, and the doubles are what
// crosses to the model code -- 8 bytes per sample.  Neither library is in this image, and the product must hand the GPU
// the STORED samples (2 bytes each) for rmg_lm_fit_stored / rmg_randomize_stored.
//
// So this file restates the container format itself, from the published HDF5 File Format Specification (versions 1.1 /
// 2.0) and the MINC 2.0 layout (/minc-2.0/image/0/{image,image-min,image-max}, /minc-2.0/dimensions/*):
//   superblock 0-3, version-1 object headers (+ continuation blocks) and version-2 ("OHDR"/"OCHK"), old-style groups
//   (symbol-table message -> version-1 B-tree of SNOD nodes + local heap) and compact new-style groups (link messages),
//   dataspace / datatype / layout (compact, contiguous, chunked with a version-1 chunk B-tree, layout versions 1-3) /
//   filter pipeline (deflate, shuffle) / attribute messages (versions 1-3).
// The file is mmap'ed; chunks are inflated by a pool of host threads straight into the caller's buffer (pinned memory or
// a column of the V x n stored-type matrix), so a 500-volume study is decoded by all cores into the layout the kernels read.
//
// Parity: the container parser is pinned to a real HDF5-library file (tests/test_minc2_reader.py reads the one HDF5 file
// in the image, written by MATLAB through libhdf5) and to files made by the test-side writer; the stored -> real
// conversion restates libminc's ((double)u - valid_min) * (max_s - min_s) / (valid_max - valid_min) + min_s, whose source is
// not in /root/reference: "parity unpinned" for that step.
#include <zlib.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rminc_b200.h"
#include "inflate_fast.hpp"

namespace {

constexpr uint64_t kUndef = ~0ull;

struct H5Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] void fail(const std::string &m) { throw H5Error(m); }

struct Datatype {
    int cls = -1;          // 0 fixed point, 1 floating point, 3 string, 9 variable length (unsupported payload)
    uint32_t size = 0;
    bool is_signed = false;
    bool big_endian = false;
};

struct Dataspace {
    int rank = 0;
    uint64_t dims[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool null_space = false;
    uint64_t count() const
    {
        if (null_space) return 0;
        uint64_t c = 1;
        for (int i = 0; i < rank; ++i) {
            if (dims[i] != 0 && c > (1ull << 48) / dims[i]) fail("dataspace larger than 2^48 elements (corrupt file?)");
            c *= dims[i];
        }
        return c;
    }
};

struct Layout {
    int cls = -1;                 // 0 compact, 1 contiguous, 2 chunked
    uint64_t addr = kUndef;       // contiguous: data; chunked: B-tree
    uint64_t size = 0;
    const uint8_t *compact = nullptr;
    int chunk_rank = 0;           // = dataset rank (the trailing element-size entry is dropped)
    uint32_t chunk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct Filter {
    int id = 0;
    std::vector<uint32_t> cd;
};

struct Attribute {
    std::string name;
    Datatype dt;
    Dataspace ds;
    const uint8_t *data = nullptr;
    size_t nbytes = 0;
};

struct Link {
    std::string name;
    uint64_t addr;
};

struct Object {
    bool has_dt = false, has_ds = false, has_layout = false, has_symtab = false;
    Datatype dt;
    Dataspace ds;
    Layout layout;
    std::vector<Filter> filters;
    std::vector<Attribute> attrs;
    std::vector<Link> links;      // new-style compact groups
    uint64_t btree = kUndef, heap = kUndef;
    bool dense_links = false;
    const Attribute *attr(const char *name) const
    {
        for (const Attribute &a : attrs)
            if (a.name == name) return &a;
        return nullptr;
    }
};

struct Chunk {
    uint64_t addr;
    uint32_t nbytes, mask;
    uint64_t off[8];
};

class H5File {
public:
    explicit H5File(const char *path)
    {
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) fail(std::string("Error opening input file: ") + path + ".");       // minc2_model :809
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size < 64) { ::close(fd_); fail(std::string(path) + ": not an HDF5 file (too small)"); }
        n_ = (size_t)st.st_size;
        void *m = mmap(nullptr, n_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) { ::close(fd_); fail(std::string(path) + ": mmap failed"); }
        p_ = static_cast<const uint8_t *>(m);
        try {
            superblock(path);
        } catch (...) {
            munmap(const_cast<uint8_t *>(p_), n_);
            ::close(fd_);
            throw;
        }
    }
    ~H5File()
    {
        if (p_) munmap(const_cast<uint8_t *>(p_), n_);
        if (fd_ >= 0) ::close(fd_);
    }
    H5File(const H5File &) = delete;
    H5File &operator=(const H5File &) = delete;

    Object root() const { return root_obj_addr_ != kUndef ? object(root_obj_addr_) : root_from_symtab(); }

    // path below the root, components separated by '/'
    bool find(const std::string &path, Object *out) const
    {
        Object cur = root();
        size_t i = 0;
        while (i < path.size()) {
            while (i < path.size() && path[i] == '/') ++i;
            size_t j = path.find('/', i);
            if (j == std::string::npos) j = path.size();
            if (j == i) break;
            uint64_t addr = kUndef;
            if (!lookup(cur, path.substr(i, j - i), &addr)) return false;
            cur = object(addr);
            i = j;
        }
        *out = cur;
        return true;
    }

    // ---------------------------------------------------------------- raw dataset bytes, file order, into dst
    void read_raw(const Object &o, uint8_t *dst, size_t dst_bytes, int n_threads) const
    {
        if (!o.has_dt || !o.has_ds || !o.has_layout) fail("object is not a dataset");
        const uint64_t total = o.ds.count() * o.dt.size;
        if (dst_bytes < total) fail("destination buffer too small for the dataset");
        if (total == 0) return;
        if (o.layout.cls == 0) {
            if (o.layout.size < total) fail("compact dataset shorter than its dataspace");
            std::memcpy(dst, o.layout.compact, total);
        } else if (o.layout.cls == 1) {
            if (o.layout.addr == kUndef) { std::memset(dst, 0, total); return; }      // never written: fill value 0
            std::memcpy(dst, at(o.layout.addr, total), total);
        } else if (o.layout.cls == 2) {
            read_chunked(o, dst, n_threads);
        } else {
            fail("unsupported data layout class");
        }
    }

private:
    // ---------------------------------------------------------------- bounds-checked access
    const uint8_t *at(uint64_t addr, uint64_t len) const
    {
        if (addr == kUndef) fail("undefined address in the file");
        const uint64_t a = addr + base_;
        if (a > n_ || len > n_ - a) fail("address beyond the end of the file (truncated or corrupt)");
        return p_ + a;
    }
    static uint64_t le(const uint8_t *p, int nbytes)
    {
        uint64_t v = 0;
        for (int i = 0; i < nbytes; ++i) v |= (uint64_t)p[i] << (8 * i);
        return v;
    }
    uint64_t off_at(const uint8_t *p) const
    {
        const uint64_t v = le(p, so_);
        return (so_ < 8 && v == ((1ull << (8 * so_)) - 1)) ? kUndef : v;
    }
    uint64_t len_at(const uint8_t *p) const { return le(p, sl_); }

    void superblock(const char *path)
    {
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        uint64_t pos = 0;
        bool found = false;
        for (; pos + 8 <= n_; pos = pos ? pos * 2 : 512) {          // the signature sits at 0, 512, 1024, 2048, ...
            if (std::memcmp(p_ + pos, sig, 8) == 0) { found = true; break; }
        }
        if (!found) fail(std::string(path) + ": no HDF5 signature (MINC 1 / NetCDF files are not MINC 2 volumes)");
        const uint8_t *s = p_ + pos;
        if (pos + 160 > n_) fail("truncated superblock");
        const int ver = s[8];
        if (ver == 0 || ver == 1) {
            so_ = s[13];
            sl_ = s[14];
            check_sizes();
            const uint8_t *q = s + 24 + (ver == 1 ? 4 : 0);
            base_ = le(q, so_);
            // free-space address, end-of-file address, driver block address follow; then the root symbol-table entry
            const uint8_t *ste = q + 4 * so_;
            root_btree_heap(ste);
        } else if (ver == 2 || ver == 3) {
            so_ = s[9];
            sl_ = s[10];
            check_sizes();
            const uint8_t *q = s + 12;
            base_ = le(q, so_);
            root_obj_addr_ = off_at(q + 3 * so_);
        } else {
            fail("unsupported HDF5 superblock version " + std::to_string(ver));
        }
        if (base_ == 0 && pos != 0) base_ = pos;      // a user block in front of the file (MATLAB writes one)
    }
    void check_sizes() const
    {
        if ((so_ != 4 && so_ != 8) || (sl_ != 4 && sl_ != 8)) fail("unsupported size of offsets / lengths");
    }
    void root_btree_heap(const uint8_t *ste)
    {
        root_header_ = off_at(ste + so_);
        const uint32_t cache = (uint32_t)le(ste + 2 * so_, 4);
        if (cache == 1) {
            root_btree_ = off_at(ste + 2 * so_ + 8);
            root_heap_ = off_at(ste + 2 * so_ + 8 + so_);
        }
    }
    Object root_from_symtab() const
    {
        Object o = object(root_header_);
        if (!o.has_symtab && root_btree_ != kUndef) {
            o.has_symtab = true;
            o.btree = root_btree_;
            o.heap = root_heap_;
        }
        return o;
    }

    // ---------------------------------------------------------------- object headers
    Object object(uint64_t addr) const
    {
        Object o;
        const uint8_t *h = at(addr, 16);
        if (std::memcmp(h, "OHDR", 4) == 0) parse_header_v2(addr, o);
        else parse_header_v1(addr, o);
        return o;
    }

    void parse_header_v1(uint64_t addr, Object &o) const
    {
        const uint8_t *h = at(addr, 16);
        if (h[0] != 1) fail("unsupported object header version " + std::to_string(h[0]));
        int nmsg = (int)le(h + 2, 2);
        const uint64_t size0 = le(h + 8, 4);
        struct Block { uint64_t addr, len; };
        std::vector<Block> blocks{{addr + 16, size0}};
        for (size_t b = 0; b < blocks.size() && nmsg > 0; ++b) {
            uint64_t pos = blocks[b].addr;
            const uint64_t end = pos + blocks[b].len;
            while (nmsg > 0 && pos + 8 <= end) {
                const uint8_t *m = at(pos, 8);
                const int type = (int)le(m, 2);
                const uint64_t sz = le(m + 2, 2);
                const int flags = m[4];
                if (pos + 8 + sz > end) fail("object header message runs past its block");
                const uint8_t *d = at(pos + 8, sz);
                if (type == 0x10) blocks.push_back({off_at(d), len_at(d + so_)});
                else message(type, flags, d, sz, o);
                pos += 8 + sz;
                --nmsg;
            }
            if (blocks.size() > 4096) fail("object header continuation loop");
        }
    }

    void parse_header_v2(uint64_t addr, Object &o) const
    {
        const uint8_t *h = at(addr, 8);
        if (h[4] != 2) fail("unsupported OHDR version");
        const int flags = h[5];
        uint64_t pos = addr + 6;
        if (flags & 0x20) pos += 16;        // four time stamps
        if (flags & 0x10) pos += 4;         // max compact / min dense attribute counts
        const int szb = 1 << (flags & 3);
        const uint64_t size0 = le(at(pos, szb), szb);
        pos += szb;
        struct Block { uint64_t addr, len; };
        std::vector<Block> blocks{{pos, size0}};
        const int mh = 4 + ((flags & 4) ? 2 : 0);
        for (size_t b = 0; b < blocks.size(); ++b) {
            uint64_t q = blocks[b].addr;
            const uint64_t end = q + blocks[b].len;          // the block's checksum follows `end`
            while (q + mh <= end) {
                const uint8_t *m = at(q, mh);
                const int type = m[0];
                const uint64_t sz = le(m + 1, 2);
                const int mflags = m[3];
                if (q + mh + sz > end) break;                // gap at the end of a block
                const uint8_t *d = at(q + mh, sz);
                if (type == 0x10) {
                    const uint64_t caddr = off_at(d), clen = len_at(d + so_);
                    if (clen < 8 || std::memcmp(at(caddr, 4), "OCHK", 4) != 0) fail("bad object header continuation block");
                    blocks.push_back({caddr + 4, clen - 8});
                } else {
                    message(type, mflags, d, sz, o);
                }
                q += mh + sz;
            }
            if (blocks.size() > 4096) fail("object header continuation loop");
        }
    }

    void message(int type, int flags, const uint8_t *d, uint64_t sz, Object &o) const
    {
        if (flags & 2) return;            // shared message: none of the MINC objects this reader needs uses them
        switch (type) {
        case 0x01: o.ds = dataspace(d, sz); o.has_ds = true; break;
        case 0x03: o.dt = datatype(d, sz); o.has_dt = true; break;
        case 0x08: o.layout = layout(d, sz); o.has_layout = true; break;
        case 0x0b: o.filters = filters(d, sz); break;
        case 0x0c: o.attrs.push_back(attribute(d, sz)); break;
        case 0x11:
            if (sz < (uint64_t)2 * so_) fail("short symbol table message");
            o.has_symtab = true;
            o.btree = off_at(d);
            o.heap = off_at(d + so_);
            break;
        case 0x02: {                      // link info: a fractal-heap address other than undefined = dense storage
            if (sz < 2) break;
            uint64_t q = 2 + ((d[1] & 1) ? 8 : 0);
            if (q + so_ <= sz && off_at(d + q) != kUndef) o.dense_links = true;
            break;
        }
        case 0x06: link(d, sz, o); break;
        default: break;                   // fill values, times, comments, ...: not needed
        }
    }

    static Dataspace with_dims(Dataspace s, const uint8_t *d, uint64_t sz, uint64_t pos, int len_bytes)
    {
        if (pos + (uint64_t)s.rank * len_bytes > sz) fail("dataspace message shorter than its rank");
        for (int i = 0; i < s.rank; ++i) s.dims[i] = le(d + pos + (uint64_t)i * len_bytes, len_bytes);
        return s;
    }
    Dataspace dataspace(const uint8_t *d, uint64_t sz) const
    {
        // the dimension sizes are "size of lengths" wide
        Dataspace s;
        if (sz < 4) fail("short dataspace message");
        const int ver = d[0];
        s.rank = d[1];
        uint64_t pos;
        if (ver == 1) pos = 8;
        else if (ver == 2) { pos = 4; if (d[3] == 2) s.null_space = true; }
        else fail("unsupported dataspace version");
        if (s.rank > 8) fail("dataspace rank above 8");
        return with_dims(s, d, sz, pos, sl_);
    }

    static Datatype datatype(const uint8_t *d, uint64_t sz)
    {
        if (sz < 8) fail("short datatype message");
        Datatype t;
        t.cls = d[0] & 0x0f;
        t.size = (uint32_t)le(d + 4, 4);
        t.big_endian = (d[1] & 1) != 0;
        if (t.cls == 0) t.is_signed = (d[1] & 8) != 0;
        return t;
    }

    Layout layout(const uint8_t *d, uint64_t sz) const
    {
        Layout l;
        if (sz < 2) fail("short layout message");
        const int ver = d[0];
        if (ver == 1 || ver == 2) {
            if (sz < 8) fail("short layout message");
            const int nd = d[1];
            l.cls = d[2];
            uint64_t pos = 8;
            if (nd > 9) fail("layout dimensionality above 9");
            if (pos + (l.cls != 0 ? so_ : 0) + 4ull * nd + (l.cls == 0 ? 4 : 0) > sz) fail("layout message shorter than its fields");
            if (l.cls != 0) { l.addr = off_at(d + pos); pos += so_; }
            uint32_t dims[9] = {0};
            for (int i = 0; i < nd; ++i, pos += 4) dims[i] = (uint32_t)le(d + pos, 4);
            if (l.cls == 2) {
                l.chunk_rank = nd - 1;
                for (int i = 0; i < nd - 1 && i < 8; ++i) l.chunk[i] = dims[i];
            } else if (l.cls == 0) {
                l.size = le(d + pos, 4);
                l.compact = d + pos + 4;
                if (pos + 4 + l.size > sz) fail("compact data runs past the layout message");
            }
        } else if (ver == 3) {
            l.cls = d[1];
            if (l.cls == 0) {
                if (sz < 4) fail("short layout message");
                l.size = le(d + 2, 2);
                l.compact = d + 4;
                if (4 + l.size > sz) fail("compact data runs past the layout message");
            } else if (l.cls == 1) {
                if (2ull + so_ + sl_ > sz) fail("layout message shorter than its fields");
                l.addr = off_at(d + 2);
                l.size = len_at(d + 2 + so_);
            } else if (l.cls == 2) {
                if (sz < 3) fail("short layout message");
                const int nd = d[2];
                if (nd < 1 || nd > 9) fail("bad chunk dimensionality");
                if (3ull + so_ + 4ull * nd > sz) fail("layout message shorter than its fields");
                l.addr = off_at(d + 3);
                l.chunk_rank = nd - 1;
                for (int i = 0; i < nd - 1; ++i) l.chunk[i] = (uint32_t)le(d + 3 + so_ + 4 * i, 4);
            } else {
                fail("unsupported layout class");
            }
        } else {
            fail("unsupported data layout version " + std::to_string(ver) + " (written with a newer HDF5 file-format bound)");
        }
        return l;
    }

    static std::vector<Filter> filters(const uint8_t *d, uint64_t sz)
    {
        std::vector<Filter> out;
        if (sz < 2) fail("short filter message");
        const int ver = d[0], nf = d[1];
        uint64_t pos = ver == 1 ? 8 : 2;
        for (int i = 0; i < nf; ++i) {
            if (pos + 8 > sz) fail("filter pipeline message too short");
            Filter f;
            f.id = (int)le(d + pos, 2);
            uint64_t namelen = 0;
            if (ver == 1 || f.id >= 256) { namelen = le(d + pos + 2, 2); pos += 4; }
            else pos += 2;
            pos += 2;                                    // flags
            const int ncd = (int)le(d + pos, 2);
            pos += 2;
            if (ver == 1) namelen = (namelen + 7) & ~7ull;
            pos += namelen;
            if (pos + 4ull * ncd > sz) fail("filter client data runs past the message");
            for (int k = 0; k < ncd; ++k) f.cd.push_back((uint32_t)le(d + pos + 4 * k, 4));
            pos += 4ull * ncd;
            if (ver == 1 && (ncd & 1)) pos += 4;
            out.push_back(f);
        }
        return out;
    }

    Attribute attribute(const uint8_t *d, uint64_t sz) const
    {
        Attribute a;
        if (sz < 8) fail("short attribute message");
        const int ver = d[0];
        const uint64_t nsz = le(d + 2, 2), tsz = le(d + 4, 2), ssz = le(d + 6, 2);
        uint64_t pos = 8;
        if (ver == 3) pos = 9;
        else if (ver != 1 && ver != 2) fail("unsupported attribute version");
        auto pad = [&](uint64_t x) { return ver == 1 ? ((x + 7) & ~7ull) : x; };
        if (pos + pad(nsz) + pad(tsz) + pad(ssz) > sz) fail("attribute message shorter than its parts");
        a.name.assign(reinterpret_cast<const char *>(d + pos), strnlen(reinterpret_cast<const char *>(d + pos), nsz));
        pos += pad(nsz);
        if ((ver >= 2) && (d[1] & 3)) fail("attribute '" + a.name + "' uses a shared datatype / dataspace");
        a.dt = datatype(d + pos, tsz);
        pos += pad(tsz);
        a.ds = dataspace(d + pos, ssz);
        pos += pad(ssz);
        a.data = d + pos;
        a.nbytes = (size_t)(sz - pos);
        if (a.dt.cls != 9 && a.ds.count() * a.dt.size > a.nbytes) fail("attribute '" + a.name + "' data runs past its message");
        return a;
    }

    void link(const uint8_t *d, uint64_t sz, Object &o) const
    {
        if (sz < 4 || d[0] != 1) return;
        const int flags = d[1];
        uint64_t pos = 2;
        int ltype = 0;
        if (flags & 8) ltype = d[pos++];
        if (flags & 4) pos += 8;
        if (flags & 16) pos += 1;
        const int lb = 1 << (flags & 3);
        if (pos + lb > sz) fail("link message shorter than its fields");
        const uint64_t nlen = le(d + pos, lb);
        pos += lb;
        if (nlen > sz || pos + nlen > sz) fail("link name runs past its message");
        std::string name(reinterpret_cast<const char *>(d + pos), nlen);
        pos += nlen;
        if (ltype == 0 && pos + so_ <= sz) o.links.push_back({name, off_at(d + pos)});
    }

    // ---------------------------------------------------------------- groups
    bool lookup(const Object &g, const std::string &name, uint64_t *addr) const
    {
        for (const Link &l : g.links)
            if (l.name == name) { *addr = l.addr; return true; }
        if (g.dense_links) fail("group uses dense link storage (fractal heap): not supported by this reader");
        if (!g.has_symtab) return false;
        const uint8_t *hp = at(g.heap, 8 + 2 * sl_ + so_);
        if (std::memcmp(hp, "HEAP", 4) != 0) fail("bad local heap signature");
        const uint64_t hsize = len_at(hp + 8);
        const uint64_t hdata = off_at(hp + 8 + 2 * sl_);
        const uint8_t *names = at(hdata, hsize);
        size_t budget = 1u << 20;
        return walk_group(g.btree, names, hsize, name, addr, 0, budget);
    }

    bool walk_group(uint64_t node, const uint8_t *names, uint64_t hsize, const std::string &name, uint64_t *addr, int depth,
                    size_t &budget) const
    {
        if (depth > 32 || budget-- == 0) fail("group B-tree too deep or cyclic");
        const uint8_t *t = at(node, 8 + 2 * so_);
        if (std::memcmp(t, "TREE", 4) == 0) {
            if (t[4] != 0) fail("group B-tree node of the wrong type");
            const int used = (int)le(t + 6, 2);
            const uint8_t *e = at(node + 8 + 2 * so_, (uint64_t)used * (sl_ + so_) + sl_);
            for (int i = 0; i < used; ++i) {
                const uint64_t child = off_at(e + sl_ + (uint64_t)i * (sl_ + so_));
                if (walk_group(child, names, hsize, name, addr, depth + 1, budget)) return true;
            }
            return false;
        }
        if (std::memcmp(t, "SNOD", 4) != 0) fail("bad symbol table node signature");
        const int nsym = (int)le(t + 6, 2);
        const uint64_t esz = 2 * so_ + 8 + 16;
        const uint8_t *e = at(node + 8, nsym * esz);
        for (int i = 0; i < nsym; ++i, e += esz) {
            const uint64_t noff = off_at(e);
            if (noff >= hsize) fail("symbol name outside the local heap");
            const char *nm = reinterpret_cast<const char *>(names + noff);
            if (name.size() == strnlen(nm, hsize - noff) && std::memcmp(nm, name.data(), name.size()) == 0) {
                *addr = off_at(e + so_);
                return true;
            }
        }
        return false;
    }

    // ---------------------------------------------------------------- chunked datasets
    void collect_chunks(uint64_t node, int rank, std::vector<Chunk> &out, int depth, size_t &budget) const
    {
        if (depth > 32 || budget-- == 0) fail("chunk B-tree too deep or cyclic");
        const uint8_t *t = at(node, 8 + 2 * so_);
        if (std::memcmp(t, "TREE", 4) != 0 || t[4] != 1) fail("bad chunk B-tree node");
        const int level = t[5];
        const int used = (int)le(t + 6, 2);
        const uint64_t ksz = 8 + 8ull * (rank + 1);
        const uint8_t *e = at(node + 8 + 2 * so_, used * (ksz + so_) + ksz);
        for (int i = 0; i < used; ++i) {
            const uint8_t *k = e + (uint64_t)i * (ksz + so_);
            const uint64_t child = off_at(k + ksz);
            if (level > 0) { collect_chunks(child, rank, out, depth + 1, budget); continue; }
            Chunk c;
            c.nbytes = (uint32_t)le(k, 4);
            c.mask = (uint32_t)le(k + 4, 4);
            for (int d = 0; d < rank; ++d) c.off[d] = le(k + 8 + 8ull * d, 8);
            c.addr = child;
            out.push_back(c);
        }
    }

    void read_chunked(const Object &o, uint8_t *dst, int n_threads) const
    {
        const int rank = o.ds.rank;
        if (o.layout.chunk_rank != rank || rank < 1) fail("chunk rank differs from the dataspace rank");
        const uint32_t es = o.dt.size;
        uint64_t cvox = 1;
        for (int d = 0; d < rank; ++d) {
            if (o.layout.chunk[d] == 0) fail("zero chunk dimension");
            cvox *= o.layout.chunk[d];
            if (cvox > (1ull << 30)) fail("chunk larger than 1 GiB");
        }
        const uint64_t cbytes = cvox * es;
        if (cbytes > (1ull << 30)) fail("chunk larger than 1 GiB");
        bool deflate = false, shuffle = false;
        for (const Filter &f : o.filters) {
            if (f.id == 1) deflate = true;
            else if (f.id == 2) shuffle = true;
            else if (f.id == 3) {}                        // fletcher32: the 4 trailing checksum bytes are ignored below
            else fail("unsupported HDF5 filter id " + std::to_string(f.id));
        }
        bool fletcher = false;
        for (const Filter &f : o.filters) fletcher |= f.id == 3;
        std::vector<Chunk> chunks;
        size_t budget = 1u << 22;
        if (o.layout.addr != kUndef) collect_chunks(o.layout.addr, rank, chunks, 0, budget);
        // voxels no chunk covers keep the fill value 0
        uint64_t covered = 0;
        for (const Chunk &c : chunks) {
            uint64_t v = 1;
            for (int d = 0; d < rank; ++d) {
                if (c.off[d] >= o.ds.dims[d]) fail("chunk offset outside the dataset");
                v *= std::min<uint64_t>(o.layout.chunk[d], o.ds.dims[d] - c.off[d]);
            }
            covered += v;
        }
        if (covered < o.ds.count()) std::memset(dst, 0, o.ds.count() * es);

        uint64_t stride[8];       // element strides of the dataset, C order
        stride[rank - 1] = 1;
        for (int d = rank - 2; d >= 0; --d) stride[d] = stride[d + 1] * o.ds.dims[d + 1];

        std::atomic<size_t> next{0};
        std::atomic<bool> bad{false};
        std::string msg;
        std::mutex mu;
        // RMG_INFLATE=zlib: zlib's uncompress for every chunk (the twin of the built-in decoder, which it also backs up)
        static const bool zlib_only = [] { const char *e = std::getenv("RMG_INFLATE"); return e && std::strcmp(e, "zlib") == 0; }();
        auto work = [&]() {
            std::vector<uint8_t> buf(cbytes), tmp(shuffle ? cbytes : 0);
            std::unique_ptr<rmg_inflate::Tables> tables(new rmg_inflate::Tables);
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= chunks.size() || bad.load()) break;
                const Chunk &c = chunks[i];
                try {
                    const uint8_t *src = at(c.addr, c.nbytes);
                    uint64_t have = c.nbytes;
                    if (fletcher && !(c.mask & (1u << filter_index(o, 3)))) { if (have < 4) fail("chunk shorter than its checksum"); have -= 4; }
                    const uint8_t *raw = src;
                    // A chunk that spans the dataset in every dimension but the slowest (libminc's slice chunks) and lies wholly
                    // inside it is one contiguous piece of the destination: inflate straight into it, no staging copy.
                    bool whole = !shuffle && c.off[0] + o.layout.chunk[0] <= o.ds.dims[0];
                    for (int d = 1; d < rank && whole; ++d) whole = c.off[d] == 0 && o.layout.chunk[d] == o.ds.dims[d];
                    if (deflate && !(c.mask & (1u << filter_index(o, 1)))) {
                        uint8_t *target = whole ? dst + c.off[0] * stride[0] * es : buf.data();
                        const int64_t got = zlib_only ? -1 : rmg_inflate::zlib_uncompress(src, (size_t)have, target, (size_t)cbytes, *tables);
                        if (got != (int64_t)cbytes) {               // not decoded (or not this decoder's day): zlib decides
                            uLongf outlen = (uLongf)cbytes;
                            const int rc = uncompress(target, &outlen, src, (uLong)have);
                            if (rc != Z_OK || outlen != cbytes) fail("inflate failed on a chunk (zlib code " + std::to_string(rc) + ")");
                        }
                        if (whole) continue;
                        raw = buf.data();
                    } else if (have < cbytes) {
                        fail("stored chunk shorter than the chunk size");
                    }
                    if (shuffle && !(c.mask & (1u << filter_index(o, 2))) && es > 1) {
                        for (uint32_t b = 0; b < es; ++b)
                            for (uint64_t v = 0; v < cvox; ++v) tmp[v * es + b] = raw[(uint64_t)b * cvox + v];
                        raw = tmp.data();
                    }
                    // copy the rows of the chunk (fastest dimension contiguous) that lie inside the dataset
                    uint64_t ext[8], idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int d = 0; d < rank; ++d) ext[d] = std::min<uint64_t>(o.layout.chunk[d], o.ds.dims[d] - c.off[d]);
                    const uint64_t rowb = ext[rank - 1] * es;
                    for (;;) {
                        uint64_t so = 0, dofs = 0, cs = 1;
                        for (int d = rank - 1; d >= 0; --d) {
                            so += idx[d] * cs;
                            cs *= o.layout.chunk[d];
                            dofs += (c.off[d] + idx[d]) * stride[d];
                        }
                        std::memcpy(dst + dofs * es, raw + so * es, rowb);
                        int d = rank - 2;
                        for (; d >= 0; --d) {
                            if (++idx[d] < ext[d]) break;
                            idx[d] = 0;
                        }
                        if (d < 0) break;
                    }
                } catch (const std::exception &e) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!bad.exchange(true)) msg = e.what();
                }
            }
        };
        const int nt = std::max(1, std::min<int>(n_threads, (int)chunks.size()));
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (std::thread &t : pool) t.join();
        if (bad.load()) fail(msg);
    }

    static int filter_index(const Object &o, int id)
    {
        for (size_t i = 0; i < o.filters.size(); ++i)
            if (o.filters[i].id == id) return (int)i;
        return 31;
    }

    int fd_ = -1;
    const uint8_t *p_ = nullptr;
    size_t n_ = 0;
    int so_ = 8, sl_ = 8;
    uint64_t base_ = 0;
    uint64_t root_header_ = kUndef, root_btree_ = kUndef, root_heap_ = kUndef, root_obj_addr_ = kUndef;
};

// ---------------------------------------------------------------- attribute values
double attr_number(const Attribute &a, size_t i)
{
    if (i >= a.ds.count() && !(a.ds.rank == 0 && i == 0)) fail("attribute '" + a.name + "' has too few values");
    const uint8_t *p = a.data + i * a.dt.size;
    uint8_t b[8];
    if (a.dt.size > 8) fail("attribute '" + a.name + "': unsupported number size");
    for (uint32_t k = 0; k < a.dt.size; ++k) b[k] = a.dt.big_endian ? p[a.dt.size - 1 - k] : p[k];
    if (a.dt.cls == 1) {
        if (a.dt.size == 8) { double v; std::memcpy(&v, b, 8); return v; }
        if (a.dt.size == 4) { float v; std::memcpy(&v, b, 4); return v; }
    } else if (a.dt.cls == 0) {
        uint64_t u = 0;
        for (uint32_t k = 0; k < a.dt.size; ++k) u |= (uint64_t)b[k] << (8 * k);
        if (a.dt.is_signed && a.dt.size < 8 && (u >> (8 * a.dt.size - 1))) u |= ~0ull << (8 * a.dt.size);
        return a.dt.is_signed ? (double)(int64_t)u : (double)u;
    }
    fail("attribute '" + a.name + "' is not numeric");
}

std::string attr_string(const Attribute &a)
{
    if (a.dt.cls != 3) fail("attribute '" + a.name + "' is not a fixed-length string");
    return std::string(reinterpret_cast<const char *>(a.data), strnlen(reinterpret_cast<const char *>(a.data), a.dt.size));
}

}  // namespace

// ------------------------------------------------------------------ the MINC 2 layer
struct rmg_minc2_volume {
    H5File file;
    Object image, image_min, image_max;
    bool has_min = false, has_max = false;
    rmg_minc2_info info;
    std::string path;
    explicit rmg_minc2_volume(const char *p) : file(p), path(p) { std::memset(&info, 0, sizeof info); }
};

namespace {

int dtype_code(const Datatype &t)
{
    if (t.cls == 0) {
        switch (t.size) {
        case 1: return t.is_signed ? RMG_MINC2_I8 : RMG_MINC2_U8;
        case 2: return t.is_signed ? RMG_MINC2_I16 : RMG_MINC2_U16;
        case 4: return t.is_signed ? RMG_MINC2_I32 : RMG_MINC2_U32;
        }
    } else if (t.cls == 1) {
        if (t.size == 4) return RMG_MINC2_F32;
        if (t.size == 8) return RMG_MINC2_F64;
    }
    fail("unsupported voxel type (class " + std::to_string(t.cls) + ", " + std::to_string(t.size) + " bytes)");
}

void default_valid_range(int code, double r[2])
{
    switch (code) {
    case RMG_MINC2_U8: r[0] = 0; r[1] = 255; break;
    case RMG_MINC2_I8: r[0] = -128; r[1] = 127; break;
    case RMG_MINC2_U16: r[0] = 0; r[1] = 65535; break;
    case RMG_MINC2_I16: r[0] = -32768; r[1] = 32767; break;
    case RMG_MINC2_U32: r[0] = 0; r[1] = 4294967295.0; break;
    case RMG_MINC2_I32: r[0] = -2147483648.0; r[1] = 2147483647.0; break;
    default: r[0] = 0; r[1] = 1; break;
    }
}

void byteswap_inplace(uint8_t *p, uint64_t count, uint32_t es)
{
    for (uint64_t i = 0; i < count; ++i, p += es)
        for (uint32_t a = 0, b = es - 1; a < b; ++a, --b) std::swap(p[a], p[b]);
}

void open_impl(rmg_minc2_volume &v)
{
    if (!v.file.find("minc-2.0/image/0/image", &v.image))
        fail(v.path + ": no /minc-2.0/image/0/image dataset (an HDF5 file, but not a MINC 2 volume)");
    if (!v.image.has_ds || !v.image.has_dt || !v.image.has_layout) fail(v.path + ": /minc-2.0/image/0/image is not a dataset");
    rmg_minc2_info &I = v.info;
    I.ndims = v.image.ds.rank;
    if (I.ndims < 1 || I.ndims > RMG_MINC2_MAX_DIMS) fail(v.path + ": image rank outside 1.." + std::to_string(RMG_MINC2_MAX_DIMS));
    I.nvoxels = 1;
    for (int d = 0; d < I.ndims; ++d) {
        I.sizes[d] = (int64_t)v.image.ds.dims[d];
        I.nvoxels *= I.sizes[d];
        I.steps[d] = 1.0;
        I.dircos[d][0] = I.dircos[d][1] = I.dircos[d][2] = 0.0;
    }
    I.dtype = dtype_code(v.image.dt);
    I.elem_size = (int32_t)v.image.dt.size;
    default_valid_range(I.dtype, I.valid_range);
    if (const Attribute *a = v.image.attr("valid_range")) {
        if (a->ds.count() >= 2) {
            const double lo = attr_number(*a, 0), hi = attr_number(*a, 1);
            I.valid_range[0] = std::min(lo, hi);
            I.valid_range[1] = std::max(lo, hi);
        }
    }
    // dimension order: the image's "dimorder" attribute, slowest first (file order)
    std::vector<std::string> names;
    if (const Attribute *a = v.image.attr("dimorder")) {
        const std::string s = attr_string(*a);
        size_t i = 0;
        while (i <= s.size()) {
            size_t j = s.find(',', i);
            if (j == std::string::npos) j = s.size();
            if (j > i) names.push_back(s.substr(i, j - i));
            i = j + 1;
        }
    }
    for (int d = 0; d < I.ndims; ++d) {
        const std::string nm = d < (int)names.size() ? names[d] : std::string();
        std::snprintf(I.dimnames[d], sizeof I.dimnames[d], "%s", nm.c_str());
        if (nm.empty()) continue;
        Object dim;
        if (!v.file.find("minc-2.0/dimensions/" + nm, &dim)) continue;
        if (const Attribute *a = dim.attr("step")) I.steps[d] = attr_number(*a, 0);
        if (const Attribute *a = dim.attr("start")) I.starts[d] = attr_number(*a, 0);
        if (const Attribute *a = dim.attr("direction_cosines"))
            for (int k = 0; k < 3 && (uint64_t)k < a->ds.count(); ++k) I.dircos[d][k] = attr_number(*a, k);
        if (const Attribute *a = dim.attr("length")) {
            if ((int64_t)attr_number(*a, 0) != I.sizes[d])
                fail(v.path + ": dimension " + nm + " has length " + std::to_string((int64_t)attr_number(*a, 0)) + " but the image has " +
                     std::to_string(I.sizes[d]));
        }
    }
    // real ranges: scalar (one range for the volume) or one entry per index of the slowest dimension(s)
    v.has_min = v.file.find("minc-2.0/image/0/image-min", &v.image_min) && v.image_min.has_ds;
    v.has_max = v.file.find("minc-2.0/image/0/image-max", &v.image_max) && v.image_max.has_ds;
    I.nslices = 1;
    if (v.has_min && v.has_max) {
        const uint64_t a = std::max<uint64_t>(v.image_min.ds.count(), 1), b = std::max<uint64_t>(v.image_max.ds.count(), 1);
        if (a != b) fail(v.path + ": image-min and image-max differ in size");
        if (v.image_min.ds.rank > I.ndims) fail(v.path + ": image-min has more dimensions than the image");
        uint64_t lead = 1;
        for (int d = 0; d < v.image_min.ds.rank; ++d) {
            if ((int64_t)v.image_min.ds.dims[d] != I.sizes[d]) fail(v.path + ": image-min does not follow the image's slowest dimensions");
            lead *= v.image_min.ds.dims[d];
        }
        I.nslices = (int64_t)(v.image_min.ds.rank == 0 ? 1 : lead);
        if (v.image_min.dt.cls != 1 || v.image_min.dt.size != 8 || v.image_max.dt.cls != 1 || v.image_max.dt.size != 8)
            fail(v.path + ": image-min / image-max are not 64-bit floats");
    }
    I.slice_len = I.nslices > 0 ? I.nvoxels / I.nslices : I.nvoxels;
    I.has_real_range = (v.has_min && v.has_max) ? 1 : 0;
}

void read_ranges_impl(const rmg_minc2_volume &v, double *lo, double *hi)
{
    const int64_t ns = v.info.nslices;
    if (!v.info.has_real_range) {
        // no real range in the file: stored values are the real values (libminc's volumes without a range)
        for (int64_t s = 0; s < ns; ++s) { lo[s] = v.info.valid_range[0]; hi[s] = v.info.valid_range[1]; }
        return;
    }
    v.file.read_raw(v.image_min, reinterpret_cast<uint8_t *>(lo), (size_t)ns * 8, 1);
    v.file.read_raw(v.image_max, reinterpret_cast<uint8_t *>(hi), (size_t)ns * 8, 1);
    if (v.image_min.dt.big_endian) byteswap_inplace(reinterpret_cast<uint8_t *>(lo), ns, 8);
    if (v.image_max.dt.big_endian) byteswap_inplace(reinterpret_cast<uint8_t *>(hi), ns, 8);
}

void read_stored_impl(const rmg_minc2_volume &v, void *dst, int64_t dst_bytes, int n_threads)
{
    const int64_t need = v.info.nvoxels * v.info.elem_size;
    if (dst_bytes < need) fail("destination holds " + std::to_string(dst_bytes) + " bytes, the volume needs " + std::to_string(need));
    v.file.read_raw(v.image, static_cast<uint8_t *>(dst), (size_t)dst_bytes, n_threads);
    if (v.image.dt.big_endian && v.info.elem_size > 1) byteswap_inplace(static_cast<uint8_t *>(dst), v.info.nvoxels, v.info.elem_size);
}

template <typename T>
void to_real(const T *u, double *y, int64_t count, int64_t slice_len, const double *lo, const double *hi, int64_t ns, double vmin,
             double vmax, bool scaled)
{
    // ((double)u - valid_min) * scale + offset with scale = (max_s - min_s) / (valid_max - valid_min), offset = min_s: each
    // operation rounded once (this translation unit is compiled without FMA contraction), the arithmetic DESIGN.md section 3
    // records for libminc's conversion and the one the device kernels reproduce
    for (int64_t i = count - 1; i >= 0; --i) {           // backwards: u may alias the front of y
        if (!scaled) { y[i] = (double)u[i]; continue; }
        const int64_t s = std::min(i / slice_len, ns - 1);
        const double scale = (hi[s] - lo[s]) / (vmax - vmin);
        const double t = ((double)u[i] - vmin) * scale;
        y[i] = t + lo[s];
    }
}

void read_real_impl(const rmg_minc2_volume &v, double *dst, int64_t count, int n_threads)
{
    const rmg_minc2_info &I = v.info;
    if (count < I.nvoxels) fail("destination holds fewer doubles than the volume has voxels");
    // the stored voxels land in the FRONT of dst and are widened from the back
    read_stored_impl(v, dst, I.nvoxels * 8, n_threads);
    std::vector<double> lo((size_t)I.nslices), hi((size_t)I.nslices);
    read_ranges_impl(v, lo.data(), hi.data());
    const bool is_float = I.dtype == RMG_MINC2_F32 || I.dtype == RMG_MINC2_F64;
    const bool scaled = !is_float && I.has_real_range && I.valid_range[1] > I.valid_range[0];
    const double vmin = I.valid_range[0], vmax = I.valid_range[1];
    switch (I.dtype) {
#define RMG_CASE(code, T) \
    case code: to_real(reinterpret_cast<const T *>(dst), dst, I.nvoxels, I.slice_len, lo.data(), hi.data(), I.nslices, vmin, vmax, scaled); break;
        RMG_CASE(RMG_MINC2_U8, uint8_t) RMG_CASE(RMG_MINC2_I8, int8_t) RMG_CASE(RMG_MINC2_U16, uint16_t)
        RMG_CASE(RMG_MINC2_I16, int16_t) RMG_CASE(RMG_MINC2_U32, uint32_t) RMG_CASE(RMG_MINC2_I32, int32_t)
        RMG_CASE(RMG_MINC2_F32, float)
#undef RMG_CASE
    case RMG_MINC2_F64: break;
    default: fail("unsupported voxel type");
    }
}

template <typename F>
int guarded(char *err, size_t errlen, F &&f)
{
    try {
        f();
        if (err && errlen) err[0] = 0;
        return RMG_OK;
    } catch (const std::exception &e) {
        if (err && errlen) std::snprintf(err, errlen, "%s", e.what());
        return RMG_ERR_IO;
    } catch (...) {
        if (err && errlen) std::snprintf(err, errlen, "unknown error in the MINC 2 reader");
        return RMG_ERR_IO;
    }
}

}  // namespace

extern "C" {

int rmg_minc2_open(const char *path, rmg_minc2_volume **vol, char *err, size_t errlen)
{
    if (!path || !vol) {
        if (err && errlen) std::snprintf(err, errlen, "rmg_minc2_open: NULL argument");
        return RMG_ERR_INVALID;
    }
    *vol = nullptr;
    return guarded(err, errlen, [&] {
        rmg_minc2_volume *v = new rmg_minc2_volume(path);
        try {
            open_impl(*v);
        } catch (...) {
            delete v;
            throw;
        }
        *vol = v;
    });
}

int rmg_minc2_get_info(const rmg_minc2_volume *vol, rmg_minc2_info *info)
{
    if (!vol || !info) return RMG_ERR_INVALID;
    *info = vol->info;
    return RMG_OK;
}

int rmg_minc2_read_ranges(const rmg_minc2_volume *vol, double *image_min, double *image_max, char *err, size_t errlen)
{
    if (!vol || !image_min || !image_max) {
        if (err && errlen) std::snprintf(err, errlen, "rmg_minc2_read_ranges: NULL argument");
        return RMG_ERR_INVALID;
    }
    return guarded(err, errlen, [&] { read_ranges_impl(*vol, image_min, image_max); });
}

int rmg_minc2_read_stored(const rmg_minc2_volume *vol, void *dst, int64_t dst_bytes, int n_threads, char *err, size_t errlen)
{
    if (!vol || !dst) {
        if (err && errlen) std::snprintf(err, errlen, "rmg_minc2_read_stored: NULL argument");
        return RMG_ERR_INVALID;
    }
    return guarded(err, errlen, [&] { read_stored_impl(*vol, dst, dst_bytes, n_threads > 0 ? n_threads : 1); });
}

int rmg_minc2_read_real(const rmg_minc2_volume *vol, double *dst, int64_t count, int n_threads, char *err, size_t errlen)
{
    if (!vol || !dst) {
        if (err && errlen) std::snprintf(err, errlen, "rmg_minc2_read_real: NULL argument");
        return RMG_ERR_INVALID;
    }
    return guarded(err, errlen, [&] { read_real_impl(*vol, dst, count, n_threads > 0 ? n_threads : 1); });
}

void rmg_minc2_close(rmg_minc2_volume *vol) { delete vol; }

int rmg_h5_read_dataset(const char *path, const char *dataset, void *dst, int64_t dst_bytes, int32_t *rank, int64_t *dims,
                        int32_t *elem_size, int32_t *type_class, int n_threads, char *err, size_t errlen)
{
    if (!path || !dataset) {
        if (err && errlen) std::snprintf(err, errlen, "rmg_h5_read_dataset: NULL argument");
        return RMG_ERR_INVALID;
    }
    return guarded(err, errlen, [&] {
        H5File f(path);
        Object o;
        if (!f.find(dataset, &o)) fail(std::string(path) + ": no object " + dataset);
        if (!o.has_ds || !o.has_dt) fail(std::string(dataset) + " is not a dataset");
        if (rank) *rank = o.ds.rank;
        if (dims) for (int d = 0; d < o.ds.rank; ++d) dims[d] = (int64_t)o.ds.dims[d];
        if (elem_size) *elem_size = (int32_t)o.dt.size;
        if (type_class) *type_class = o.dt.cls == 0 ? (o.dt.is_signed ? 2 : 1) : (o.dt.cls == 1 ? 3 : (o.dt.cls == 3 ? 4 : 0));
        if (dst) {
            f.read_raw(o, static_cast<uint8_t *>(dst), (size_t)dst_bytes, n_threads > 0 ? n_threads : 1);
            if (o.dt.big_endian && o.dt.size > 1 && o.dt.cls <= 1) byteswap_inplace(static_cast<uint8_t *>(dst), o.ds.count(), o.dt.size);
        }
    });
}

int rmg_h5_read_attribute(const char *path, const char *object, const char *attribute, double *numbers, int64_t max_numbers,
                          int64_t *count, char *text, size_t textlen, char *err, size_t errlen)
{
    if (!path || !object || !attribute) {
        if (err && errlen) std::snprintf(err, errlen, "rmg_h5_read_attribute: NULL argument");
        return RMG_ERR_INVALID;
    }
    return guarded(err, errlen, [&] {
        H5File f(path);
        Object o;
        if (!f.find(object, &o)) fail(std::string(path) + ": no object " + object);
        const Attribute *a = o.attr(attribute);
        if (!a) fail(std::string(object) + " has no attribute " + attribute);
        if (count) *count = 0;
        if (text && textlen) text[0] = 0;
        if (a->dt.cls == 3) {
            if (text && textlen) std::snprintf(text, textlen, "%s", attr_string(*a).c_str());
            if (count) *count = 1;
        } else {
            const int64_t c = (int64_t)std::max<uint64_t>(a->ds.count(), a->ds.rank == 0 ? 1 : 0);
            if (count) *count = c;
            for (int64_t i = 0; numbers && i < c && i < max_numbers; ++i) numbers[i] = attr_number(*a, (size_t)i);
        }
    });
}

int64_t rmg_zlib_uncompress(const void *src, int64_t src_bytes, void *dst, int64_t dst_bytes, int use_zlib)
{
    if (!src || !dst || src_bytes < 0 || dst_bytes < 0) return -1;
    if (use_zlib) {
        uLongf outlen = (uLongf)dst_bytes;
        const int rc = uncompress(static_cast<Bytef *>(dst), &outlen, static_cast<const Bytef *>(src), (uLong)src_bytes);
        return rc == Z_OK ? (int64_t)outlen : -1;
    }
    std::unique_ptr<rmg_inflate::Tables> tables(new rmg_inflate::Tables);
    return rmg_inflate::zlib_uncompress(static_cast<const uint8_t *>(src), (size_t)src_bytes, static_cast<uint8_t *>(dst), (size_t)dst_bytes,
                                        *tables);
}

int rmg_minc2_read_table(const char *const *paths, int n, void *U, int64_t ldU, int dtype, double *image_min, double *image_max,
                         int64_t nslices, double *valid_range, int n_threads, char *err, size_t errlen)
{
    if (!paths || n <= 0 || !U) {
        if (err && errlen) std::snprintf(err, errlen, "rmg_minc2_read_table: NULL argument or n <= 0");
        return RMG_ERR_INVALID;
    }
    return guarded(err, errlen, [&] {
        const int es = dtype == RMG_MINC2_U16 ? 2 : (dtype == RMG_MINC2_F32 ? 4 : (dtype == RMG_MINC2_F64 ? 8 : 0));
        if (!es) fail("rmg_minc2_read_table: the table's type must be uint16, float32 or float64");
        std::atomic<int> next{0};
        std::atomic<bool> bad{false};
        std::string msg;
        std::mutex mu;
        int64_t V0 = -1;
        double vr0[2] = {0, 0};
        auto work = [&]() {
            for (;;) {
                const int j = next.fetch_add(1);
                if (j >= n || bad.load()) break;
                try {
                    if (!paths[j]) fail("rmg_minc2_read_table: NULL file name");
                    rmg_minc2_volume v(paths[j]);
                    open_impl(v);
                    const rmg_minc2_info &I = v.info;
                    if (dtype != RMG_MINC2_F64 && I.dtype != dtype) fail(v.path + ": stored type differs from the table's type");
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        if (V0 < 0) { V0 = I.nvoxels; vr0[0] = I.valid_range[0]; vr0[1] = I.valid_range[1]; }
                        if (I.nvoxels != V0) fail(v.path + ": volume sizes differ");          // minc2_model's dimension check
                        if (dtype != RMG_MINC2_F64 && (I.valid_range[0] != vr0[0] || I.valid_range[1] != vr0[1]))
                            fail(v.path + ": volumes do not share one valid range");
                    }
                    if (I.nvoxels > ldU) fail("rmg_minc2_read_table: ldU is smaller than the volumes");
                    uint8_t *col = static_cast<uint8_t *>(U) + (size_t)j * ldU * es;
                    if (dtype == RMG_MINC2_F64) {
                        read_real_impl(v, reinterpret_cast<double *>(col), I.nvoxels, 1);
                    } else {
                        read_stored_impl(v, col, I.nvoxels * es, 1);
                        if (image_min && image_max) {
                            if (I.nslices != nslices && I.nslices != 1) fail(v.path + ": volumes disagree on the number of range entries");
                            std::vector<double> lo((size_t)I.nslices), hi((size_t)I.nslices);
                            read_ranges_impl(v, lo.data(), hi.data());
                            for (int64_t s = 0; s < nslices; ++s) {
                                image_min[s + (int64_t)j * nslices] = lo[I.nslices == 1 ? 0 : s];
                                image_max[s + (int64_t)j * nslices] = hi[I.nslices == 1 ? 0 : s];
                            }
                        }
                    }
                } catch (const std::exception &e) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!bad.exchange(true)) msg = e.what();
                }
            }
        };
        const int nt = std::max(1, std::min(n_threads, n));
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (std::thread &t : pool) t.join();
        if (bad.load()) fail(msg);
        if (valid_range) { valid_range[0] = vr0[0]; valid_range[1] = vr0[1]; }
    });
}

}  // extern "C"
