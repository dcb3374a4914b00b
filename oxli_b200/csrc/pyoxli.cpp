// pyoxli.cpp -- host-side mirror of the reference's pyo3 class.
//
// The reference exposes one class, `oxli.KmerCountTable`
// (/root/reference/src/lib.rs:29-838), from a Rust crate.  No Rust toolchain
// exists in this image, so the mirror is C++/pybind11; it keeps the reference's
// method names, argument meaning, return types and exception types, and holds
// no counting logic of its own: every hash and every count comes from the CUDA
// library through the C ABI in include/oxli_b200.h.  Host-only state is what
// the reference also keeps outside the hash map: `consumed`, `version`,
// `store_kmers` and the optional hash -> k-mer string map.
//
// Supersets (keyword-only / extra methods, defaults keep reference behaviour):
//   KmerCountTable(ksize, store_kmers=False, *, device=0, capacity_hint=0)
//   consume_many(seqs, skip_bad_kmers=True)      one GPU batch for many reads
//   consume_buffer(bases, offsets, skip_bad_kmers=True)   CSR batch, zero-copy from buffers
//   KmerCountTable(..., deferred=True)           the default: consume(seq) parks reads in a pinned
//                                                 batch and returns at host speed (see Table::sync);
//                                                 deferred=False or OXLI_B200_DEFERRED=0: one
//                                                 synchronous GPU call per consume(seq)
//   consume_file(path, skip_bad_kmers=True)      FASTA/FASTQ (plain or gzip) parsed into pinned
//                                                 batches -- the loop the reference leaves to
//                                                 screed (README.md:89-98)
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "oxli_b200.h"

namespace py = pybind11;

namespace {

[[noreturn]] void raise_status(oxg_status st) {
    const std::string msg = oxg_last_error();
    switch (st) {
    case OXG_ERR_BAD_KMER:
    case OXG_ERR_INVALID: throw py::value_error(msg);
    case OXG_ERR_WRONG_KSIZE: throw py::value_error(msg);
    case OXG_ERR_NOMEM: PyErr_SetString(PyExc_MemoryError, msg.c_str()); throw py::error_already_set();
    default: throw std::runtime_error("oxli_b200: " + msg);
    }
}
inline void ck(oxg_status st) { if (st != OXG_OK) raise_status(st); }

[[noreturn]] void raise_os_error(const std::string &path) {
    PyErr_SetFromErrnoWithFilename(PyExc_OSError, path.c_str());
    throw py::error_already_set();
}

std::string upper(const std::string &s) {
    std::string u = s;
    for (auto &c : u) if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return u;
}
bool is_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
char comp(char c) { return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c; }
std::string revcomp(const std::string &s) {
    std::string r(s.rbegin(), s.rend());
    for (auto &c : r) c = comp(c);
    return r;
}

// ---- the tiny JSON subset serde_json writes for this struct (src/lib.rs:29-39) ----
struct JsonIn {
    const std::string &s;
    size_t i = 0;
    explicit JsonIn(const std::string &str) : s(str) {}
    [[noreturn]] void bad(const char *what) const {
        throw std::runtime_error(std::string("Deserialization error: ") + what + " at offset " + std::to_string(i));
    }
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    void need(char c) { if (!eat(c)) bad("unexpected character"); }
    std::string str() {
        need('"');
        std::string out;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\') { ++i; if (i >= s.size()) bad("bad escape"); }
            out.push_back(s[i++]);
        }
        if (i >= s.size()) bad("unterminated string");
        ++i;
        return out;
    }
    uint64_t u64() {
        ws();
        if (i >= s.size() || s[i] < '0' || s[i] > '9') bad("expected a number");
        uint64_t v = 0;
        while (i < s.size() && s[i] >= '0' && s[i] <= '9') v = v * 10 + (uint64_t)(s[i++] - '0');
        return v;
    }
    bool lit(const char *w) {
        ws();
        const size_t n = strlen(w);
        if (s.compare(i, n, w) == 0) { i += n; return true; }
        return false;
    }
};

// FASTA / FASTQ reader feeding pinned CSR batches to the GPU.  Records are what
// screed would hand to consume(): FASTA sequence lines joined, FASTQ 4-line records.
struct PinnedBatch {
    uint8_t *bases = nullptr;
    uint64_t cap = 0, used = 0;
    std::vector<uint64_t> offs{0};
    explicit PinnedBatch(uint64_t bytes) : cap(bytes) {
        void *p = nullptr;
        ck(oxg_pinned_alloc(bytes, &p));
        bases = static_cast<uint8_t *>(p);
    }
    ~PinnedBatch() { oxg_pinned_free(bases); }
    void grow(uint64_t need) {
        uint64_t ncap = cap;
        while (ncap < need) ncap *= 2;
        void *p = nullptr;
        ck(oxg_pinned_alloc(ncap, &p));
        memcpy(p, bases, used);
        oxg_pinned_free(bases);
        bases = static_cast<uint8_t *>(p);
        cap = ncap;
    }
    void append(const char *d, size_t n) {
        if (used + n > cap) grow(used + n);
        memcpy(bases + used, d, n);
        used += n;
    }
    void end_record() { offs.push_back(used); }
    uint64_t records() const { return offs.size() - 1; }
    void reset() { used = 0; offs.assign(1, 0); }
};

struct Table {
    oxg_table *h = nullptr;
    // deferred=True: consume(seq) in skip mode only appends the read to a pinned batch and
    // returns the number of countable windows found by a host scan; the batch goes to the
    // GPU when it fills up or when any other method needs the table (sync()).
    bool deferred = false;
    mutable std::unique_ptr<PinnedBatch> pending;
    static constexpr uint64_t kPendingBytes = 64ull << 20;
    uint8_t ksize = 0;
    bool store_kmers = false;
    uint64_t consumed = 0;
    std::string version;
    std::unordered_map<uint64_t, std::string> hash_to_kmer;
    int device = 0;

    Table(uint8_t k, bool store, int dev, uint64_t hint, bool defer = false)
        : deferred(defer), ksize(k), store_kmers(store), version(oxg_version()), device(dev) {
        ck(oxg_table_create(dev, k, hint, &h));
    }

    // push reads parked by deferred consume() calls to the GPU
    void sync() const {
        if (!pending || pending->records() == 0) return;
        // the batch leaves `pending` before the GIL is released: a consume() from another
        // Python thread meanwhile starts a fresh one instead of appending to memory in flight
        std::unique_ptr<PinnedBatch> batch = std::move(pending);
        uint64_t n = 0, ep = 0;
        int64_t er = -1;
        oxg_status st;
        {
            py::gil_scoped_release nogil;
            st = oxg_consume_batch(h, batch->bases, batch->offs.data(), batch->records(), 1, &n, &er, &ep);
        }
        batch->reset();
        if (!pending) pending = std::move(batch);
        ck(st);
    }
    ~Table() { if (h) oxg_table_destroy(h); }
    Table(const Table &) = delete;
    Table &operator=(const Table &) = delete;

    // one hash per window of `seq`; 0 marks a bad window (sourmash force=true)
    std::vector<uint64_t> window_hashes(const std::string &seq) const {
        sync();
        std::vector<uint64_t> out(seq.size() >= ksize ? seq.size() - ksize + 1 : 0);
        if (!out.empty()) ck(oxg_hash_windows(h, reinterpret_cast<const uint8_t *>(seq.data()), seq.size(), out.data()));
        return out;
    }

    // src/lib.rs:65-81
    uint64_t hash_kmer(const std::string &kmer) const {
        if ((uint8_t)kmer.size() != ksize) throw std::runtime_error("wrong ksize");
        auto hs = window_hashes(kmer);
        if (hs.empty() || hs[0] == 0) throw std::runtime_error("invalid DNA character in input k-mer: " + kmer);
        return hs[0];
    }

    // src/lib.rs:107-142
    std::string canon(const std::string &kmer) const {
        if (kmer.size() != (size_t)ksize) throw py::value_error("kmer size does not match count table ksize");
        std::string u = upper(kmer);
        for (char c : u) if (!is_acgt(c)) throw py::value_error("kmer contains invalid characters");
        std::string r = revcomp(u);
        return u <= r ? u : r;
    }

    uint64_t count_hash(uint64_t hv) {
        sync();
        uint64_t now = 0;
        ck(oxg_count_hashes(h, &hv, 1, &now));
        return now;
    }

    // src/lib.rs:145-167
    uint64_t count(const std::string &kmer) {
        if ((uint8_t)kmer.size() != ksize) throw py::value_error("kmer size does not match count table ksize");
        const uint64_t hv = hash_kmer(kmer);
        const uint64_t c = count_hash(hv);
        consumed += kmer.size();
        if (store_kmers) hash_to_kmer[hv] = canon(kmer);
        return c;
    }

    uint64_t get_hash(uint64_t hv) const {
        sync();
        uint64_t c = 0;
        ck(oxg_get_hashes(h, &hv, 1, &c));
        return c;
    }

    // src/lib.rs:170-182 (the reference panics on a non-ACGT k-mer here; a RuntimeError is raised instead)
    uint64_t get(const std::string &kmer) const {
        if ((uint8_t)kmer.size() != ksize) throw py::value_error("kmer size does not match count table ksize");
        return get_hash(hash_kmer(kmer));
    }

    std::vector<uint64_t> get_hash_array(const std::vector<uint64_t> &keys) const {
        sync();
        std::vector<uint64_t> out(keys.size());
        if (!keys.empty()) ck(oxg_get_hashes(h, keys.data(), keys.size(), out.data()));
        return out;
    }

    void drop_hash(uint64_t hv) {
        sync();
        ck(oxg_erase_hashes(h, &hv, 1, nullptr));
    }

    uint64_t cut(int mode, uint64_t thresh) {
        sync();
        uint64_t n = 0;
        ck(oxg_cut(h, mode, thresh, &n));
        return n;
    }

    uint64_t len() const {
        sync();
        uint64_t n = 0;
        ck(oxg_table_len(h, &n));
        return n;
    }

    oxg_stats stats() const {
        sync();
        oxg_stats s{};
        ck(oxg_table_stats(h, &s));
        return s;
    }

    std::vector<std::pair<uint64_t, uint64_t>> items(int sort_mode) const {
        sync();
        uint64_t n = 0;
        ck(oxg_export(h, nullptr, nullptr, 0, sort_mode, &n));
        std::vector<uint64_t> k(n + 1), v(n + 1);
        if (n) ck(oxg_export(h, k.data(), v.data(), n, sort_mode, &n));
        std::vector<std::pair<uint64_t, uint64_t>> out(n);
        for (uint64_t i = 0; i < n; ++i) out[i] = {k[i], v[i]};
        return out;
    }

    // src/lib.rs:545-607
    uint64_t consume(const std::string &seq, bool skip_bad) {
        uint64_t n = 0;
        if (store_kmers) {
            // KmersAndHashesIter path (src/lib.rs:552-573): hashes come from the GPU, the
            // k-mer strings for the side map are cut from the sequence here
            const std::string up = upper(seq);
            const auto hs = window_hashes(up);
            std::vector<uint64_t> good;
            good.reserve(hs.size());
            for (size_t i = 0; i < hs.size(); ++i) {
                const std::string sub = up.substr(i, ksize);
                if (hs[i] == 0) {
                    fprintf(stderr, "bad k-mer at position %zu: %s\n", i + 1, sub.c_str());
                    continue;
                }
                const std::string rc = revcomp(sub);
                hash_to_kmer[hs[i]] = sub < rc ? sub : rc;
                good.push_back(hs[i]);
            }
            if (!good.empty()) ck(oxg_count_hashes(h, good.data(), good.size(), nullptr));
            n = good.size();
        } else if (deferred) {
            // The per-record loop of the reference's README (`for record in ...: t.consume(seq)`)
            // must not pay a kernel launch per 150-bp read: the read is parked in a pinned batch
            // that goes to the GPU when it fills up or when any other method needs the table.
            // What the call returns comes from a host scan: windows without a non-ACGT byte --
            // exactly what the device counts, except for the reference's hash==0 skip
            // (src/lib.rs:589; probability 2^-64 per k-mer), which the scan cannot see.
            if (!pending) pending = std::make_unique<PinnedBatch>(1 << 20);  // doubles as reads arrive, flushed at kPendingBytes
            int64_t last_bad = -1, first_bad = -1;
            for (size_t i = 0; i < seq.size(); ++i) {
                const char c = seq[i] & ~0x20;
                if (!is_acgt(c)) { last_bad = (int64_t)i; if (first_bad < 0) first_bad = (int64_t)i; }
                if (i + 1 >= ksize && last_bad < (int64_t)(i + 1 - ksize)) ++n;
            }
            if (!skip_bad && first_bad >= 0 && seq.size() >= ksize) {
                // error mode (src/lib.rs:593-596): the windows before the first bad one are counted
                // and stay counted, then ValueError; `consumed` is not updated on this path
                const uint64_t w = first_bad + 1 >= (int64_t)ksize ? (uint64_t)(first_bad + 1 - ksize) : 0;
                if (w > 0) {
                    pending->append(seq.data(), w + ksize - 1);
                    pending->end_record();
                    if (pending->used >= kPendingBytes) sync();
                }
                throw py::value_error("bad k-mer encountered at position " + std::to_string(w));
            }
            pending->append(seq.data(), seq.size());
            pending->end_record();
            if (pending->used >= kPendingBytes) sync();
        } else {
            sync();
            const uint64_t offs[2] = {0, seq.size()};
            int64_t er = -1;
            uint64_t ep = 0;
            const oxg_status st = oxg_consume_batch(h, reinterpret_cast<const uint8_t *>(seq.data()), offs, 1,
                                                    skip_bad ? 1 : 0, &n, &er, &ep);
            if (st == OXG_ERR_BAD_KMER)  // early return: `consumed` is not updated (src/lib.rs:593-596)
                throw py::value_error("bad k-mer encountered at position " + std::to_string(ep));
            ck(st);
        }
        consumed += seq.size();
        return n;
    }

    uint64_t consume_csr(const uint8_t *bases, const uint64_t *offs, uint64_t n_reads, bool skip_bad) {
        sync();
        if (store_kmers) throw py::value_error("batch ingest is not available when store_kmers=True");
        uint64_t n = 0, ep = 0;
        int64_t er = -1;
        oxg_status st;
        {
            py::gil_scoped_release nogil;
            st = oxg_consume_batch(h, bases, offs, n_reads, skip_bad ? 1 : 0, &n, &er, &ep);
        }
        if (st == OXG_ERR_BAD_KMER) {
            consumed += offs[er] - offs[0];  // the reads before the failing one went through consume() in full
            throw py::value_error("bad k-mer encountered at position " + std::to_string(ep) + " (read " + std::to_string(er) + ")");
        }
        ck(st);
        consumed += offs[n_reads] - offs[0];
        return n;
    }

    // src/lib.rs:683-703 + 853-950
    std::vector<std::pair<std::string, uint64_t>> kmers_and_hashes(const std::string &seq, bool skip_bad) const {
        std::vector<std::pair<std::string, uint64_t>> out;
        const std::string up = upper(seq);
        const auto hs = window_hashes(up);
        for (size_t i = 0; i < hs.size(); ++i) {
            const std::string sub = up.substr(i, ksize);
            if (hs[i] != 0) {
                const std::string rc = revcomp(sub);
                out.emplace_back(sub < rc ? sub : rc, hs[i]);
            } else {
                fprintf(stderr, "bad k-mer at position %zu: %s\n", i + 1, sub.c_str());
                if (!skip_bad) out.emplace_back("", 0);
            }
        }
        return out;
    }

    std::vector<uint64_t> setop(const Table &o, int op) const {
        sync(); o.sync();
        uint64_t n = 0;
        const uint64_t cap = len() + o.len() + 2;
        std::vector<uint64_t> out(cap);
        ck(oxg_setop_export(h, o.h, op, out.data(), cap, &n));
        out.resize(n);
        return out;
    }

    // src/lib.rs:270-272: field order and spelling of serde_json
    std::string serialize_json() const {
        std::string s = "{\"counts\":{";
        bool first = true;
        for (auto &kv : items(0)) {
            if (!first) s += ',';
            first = false;
            s += '"' + std::to_string(kv.first) + "\":" + std::to_string(kv.second);
        }
        s += "},\"ksize\":" + std::to_string((unsigned)ksize) + ",\"version\":\"" + version + "\",\"consumed\":" +
             std::to_string(consumed) + ",\"store_kmers\":" + (store_kmers ? "true" : "false") + ",\"hash_to_kmer\":";
        if (!store_kmers) s += "null";
        else {
            s += '{';
            first = true;
            for (auto &kv : hash_to_kmer) {
                if (!first) s += ',';
                first = false;
                s += '"' + std::to_string(kv.first) + "\":\"" + kv.second + '"';
            }
            s += '}';
        }
        s += '}';
        return s;
    }
};

py::tuple consume_file(Table &t, const std::string &path, bool skip_bad, uint64_t batch_bytes) {
    FILE *probe = fopen(path.c_str(), "rb");
    if (!probe) raise_os_error(path);
    fclose(probe);
    gzFile gz = gzopen(path.c_str(), "rb");
    if (!gz) raise_os_error(path);
    gzbuffer(gz, 1 << 20);
    PinnedBatch batch(4 << 20);  // grows by doubling towards batch_bytes: a small file should not pin 256 MiB
    uint64_t n_records = 0, n_kmers = 0;
    auto flush = [&]() {
        if (batch.records() == 0) return;
        uint64_t got = 0;
        try {
            got = t.consume_csr(batch.bases, batch.offs.data(), batch.records(), skip_bad);
        } catch (...) { gzclose(gz); throw; }
        n_kmers += got;
        n_records += batch.records();
        batch.reset();
    };
    // Line-oriented state machine; gzgets may hand a long line over in several pieces.
    enum Kind { kNone, kFastaHeader, kFastaSeq, kFastqHeader, kFastqSeq, kFastqPlus, kFastqQual };
    std::vector<char> piece(1 << 16);
    int state = 0;  // 0: between records; 1: inside a FASTA record; 2/3/4: FASTQ sequence / '+' / quality line next
    Kind kind = kNone;
    bool at_line_start = true, in_record = false;
    while (gzgets(gz, piece.data(), (int)piece.size())) {
        size_t n = strlen(piece.data());
        const bool complete = n && piece[n - 1] == '\n';
        while (n && (piece[n - 1] == '\n' || piece[n - 1] == '\r')) --n;
        const char *d = piece.data();
        if (at_line_start) {
            if (state <= 1 && n && d[0] == '>') {
                if (in_record) { batch.end_record(); if (batch.used >= batch_bytes) flush(); }
                in_record = true;
                kind = kFastaHeader;
            } else if (state == 1) kind = kFastaSeq;
            else if (state == 0 && n && d[0] == '@') { in_record = true; kind = kFastqHeader; }
            else if (state == 2) kind = kFastqSeq;
            else if (state == 3) kind = kFastqPlus;
            else if (state == 4) kind = kFastqQual;
            else if (n == 0) kind = kNone;  // blank line between records
            else { gzclose(gz); throw py::value_error("not a FASTA/FASTQ file: " + path); }
        }
        if (kind == kFastaSeq || kind == kFastqSeq) batch.append(d, n);
        if (complete) {
            switch (kind) {
            case kFastaHeader: case kFastaSeq: state = 1; break;
            case kFastqHeader: state = 2; break;
            case kFastqSeq: state = 3; break;
            case kFastqPlus: state = 4; break;
            case kFastqQual:
                batch.end_record(); in_record = false; state = 0;
                if (batch.used >= batch_bytes) flush();
                break;
            default: break;
            }
        }
        at_line_start = complete;
    }
    if (in_record) batch.end_record();  // last FASTA record / truncated FASTQ record
    flush();
    gzclose(gz);
    return py::make_tuple(n_records, n_kmers);
}

py::set to_pyset(const std::vector<uint64_t> &v) {
    py::set s;
    for (uint64_t x : v) s.add(py::int_(x));
    return s;
}

void write_pairs_tsv(const std::string &path, const std::vector<std::pair<std::string, uint64_t>> &rows) {
    FILE *f = fopen(path.c_str(), "w");
    if (!f) raise_os_error(path);
    for (auto &r : rows) fprintf(f, "%s\t%llu\n", r.first.c_str(), (unsigned long long)r.second);
    fclose(f);
}

// src/lib.rs:275-293: gzip level 1 (niffler Level::One)
void save(const Table &t, const std::string &path) {
    FILE *probe = fopen(path.c_str(), "wb");
    if (!probe) raise_os_error(path);
    fclose(probe);
    gzFile gz = gzopen(path.c_str(), "wb1");
    if (!gz) raise_os_error(path);
    const std::string js = t.serialize_json();
    size_t off = 0;
    while (off < js.size()) {
        const unsigned chunk = (unsigned)std::min<size_t>(js.size() - off, 1u << 30);
        if (gzwrite(gz, js.data() + off, chunk) <= 0) { gzclose(gz); errno = EIO; raise_os_error(path); }
        off += chunk;
    }
    gzclose(gz);
}

// src/lib.rs:295-322: gzip is auto-detected (zlib reads plain files transparently)
std::unique_ptr<Table> load(const std::string &path, int device) {
    FILE *probe = fopen(path.c_str(), "rb");
    if (!probe) raise_os_error(path);
    fclose(probe);
    gzFile gz = gzopen(path.c_str(), "rb");
    if (!gz) raise_os_error(path);
    std::string js;
    char buf[1 << 16];
    int n;
    while ((n = gzread(gz, buf, sizeof buf)) > 0) js.append(buf, (size_t)n);
    gzclose(gz);

    JsonIn in(js);
    std::vector<uint64_t> keys, vals;
    std::unordered_map<uint64_t, std::string> h2k;
    uint64_t ksize = 0, consumed = 0;
    bool store = false, have_counts = false, have_ksize = false, have_version = false;
    std::string version;
    if (!in.eat('{')) in.bad("expected value");
    if (!in.eat('}')) {
        do {
            const std::string key = in.str();
            in.need(':');
            if (key == "counts") {
                have_counts = true;
                in.need('{');
                if (!in.eat('}')) {
                    do {
                        const std::string hk = in.str();
                        in.need(':');
                        keys.push_back(std::stoull(hk));
                        vals.push_back(in.u64());
                    } while (in.eat(','));
                    in.need('}');
                }
            } else if (key == "ksize") { ksize = in.u64(); have_ksize = true; }
            else if (key == "version") { version = in.str(); have_version = true; }
            else if (key == "consumed") consumed = in.u64();
            else if (key == "store_kmers") { if (in.lit("true")) store = true; else if (in.lit("false")) store = false; else in.bad("expected a boolean"); }
            else if (key == "hash_to_kmer") {
                if (!in.lit("null")) {
                    in.need('{');
                    if (!in.eat('}')) {
                        do {
                            const std::string hk = in.str();
                            in.need(':');
                            h2k[std::stoull(hk)] = in.str();
                        } while (in.eat(','));
                        in.need('}');
                    }
                }
            } else in.bad("unknown field");
        } while (in.eat(','));
        in.need('}');
    }
    if (!have_counts) throw std::runtime_error("Deserialization error: missing field `counts`");
    if (!have_ksize) throw std::runtime_error("Deserialization error: missing field `ksize`");
    if (!have_version) throw std::runtime_error("Deserialization error: missing field `version`");
    if (ksize > 255) throw std::runtime_error("Deserialization error: invalid value: ksize out of range for u8");
    auto t = std::make_unique<Table>((uint8_t)ksize, store, device, keys.size());
    if (!keys.empty()) ck(oxg_add_pairs(t->h, keys.data(), vals.data(), keys.size()));
    t->consumed = consumed;
    t->hash_to_kmer = std::move(h2k);
    if (version != t->version)  // src/lib.rs:314-319
        fprintf(stderr, "Version mismatch: loaded version is %s, but current version is %s\n", version.c_str(), t->version.c_str());
    t->version = version;
    return t;
}

uint8_t ksize_from_py(const py::object &o) {
    // pyo3 extracts `u8`: out-of-range ints raise OverflowError
    const long long v = o.cast<long long>();
    if (v < 0 || v > 255) { PyErr_SetString(PyExc_OverflowError, "out of range integral type conversion attempted"); throw py::error_already_set(); }
    if (v == 0) throw py::value_error("ksize must be at least 1");
    return (uint8_t)v;
}

}  // namespace

PYBIND11_MODULE(_oxli, m) {
    m.doc() = "B200-native drop-in for oxli.KmerCountTable (CUDA, sm_100a; no CPU fallback)";
    m.attr("__version__") = oxg_version();
    m.def("device_count", []() { return oxg_device_count(); });

    py::class_<Table>(m, "KmerCountTable")
        .def(py::init([](py::object ksize, bool store_kmers, int device, uint64_t capacity_hint, py::object deferred) {
                 // deferred=None (default): parked consume, unless OXLI_B200_DEFERRED=0 asks for one launch per call
                 bool defer = true;
                 if (deferred.is_none()) {
                     const char *e = getenv("OXLI_B200_DEFERRED");
                     if (e && *e) defer = strcmp(e, "0") != 0;
                 } else defer = deferred.cast<bool>();
                 return std::make_unique<Table>(ksize_from_py(ksize), store_kmers, device, capacity_hint, defer);
             }),
             py::arg("ksize"), py::arg("store_kmers") = false, py::kw_only(), py::arg("device") = 0,
             py::arg("capacity_hint") = 0, py::arg("deferred") = py::none())
        .def_property_readonly("deferred", [](const Table &t) { return t.deferred; })
        .def("flush", &Table::sync, "send reads parked by deferred consume() calls to the GPU")
        .def("hash_kmer", &Table::hash_kmer, py::arg("kmer"))
        .def("unhash", [](const Table &t, uint64_t hv) {
                 if (!t.store_kmers) throw py::value_error("K-mer storage is not enabled.");
                 auto it = t.hash_to_kmer.find(hv);
                 if (it == t.hash_to_kmer.end())
                     throw py::key_error("Warning: Hash " + std::to_string(hv) + " not found in table.");
                 return it->second;
             }, py::arg("hash"))
        .def("count_hash", &Table::count_hash, py::arg("hashval"))
        .def("canon", &Table::canon, py::arg("kmer"))
        .def("count", &Table::count, py::arg("kmer"))
        .def("get", &Table::get, py::arg("kmer"))
        .def("get_hash", &Table::get_hash, py::arg("hashval"))
        .def("get_hash_array", &Table::get_hash_array, py::arg("hash_keys"))
        .def("drop", [](Table &t, const std::string &kmer) { t.drop_hash(t.hash_kmer(kmer)); }, py::arg("kmer"))
        .def("drop_hash", &Table::drop_hash, py::arg("hashval"))
        .def("mincut", [](Table &t, uint64_t m) { return t.cut(0, m); }, py::arg("min_count"))
        .def("maxcut", [](Table &t, uint64_t m) { return t.cut(1, m); }, py::arg("max_count"))
        .def("serialize_json", &Table::serialize_json)
        .def("save", [](const Table &t, const std::string &path) { save(t, path); }, py::arg("filepath"))
        .def_static("load", [](const std::string &path, int device) { return load(path, device); },
                    py::arg("filepath"), py::kw_only(), py::arg("device") = 0)
        .def("dump", [](const Table &t, py::object file, bool sortcounts, bool sortkeys) {
                 if (sortcounts && sortkeys) throw py::value_error("Cannot sort by both counts and keys at the same time.");
                 auto rows = t.items(sortkeys ? 1 : sortcounts ? 2 : 0);
                 if (!file.is_none()) {
                     const std::string path = file.cast<std::string>();
                     FILE *f = fopen(path.c_str(), "w");
                     if (!f) raise_os_error(path);
                     for (auto &r : rows) fprintf(f, "%llu\t%llu\n", (unsigned long long)r.first, (unsigned long long)r.second);
                     fclose(f);
                     rows.clear();
                 }
                 return rows;
             }, py::arg("file") = py::none(), py::arg("sortcounts") = false, py::arg("sortkeys") = false)
        .def("dump_kmers", [](const Table &t, py::object file, bool sortcounts, bool sortkeys) {
                 if (!t.store_kmers) throw py::value_error("K-mer storage is disabled. No hash:kmer map is available.");
                 if (sortcounts && sortkeys) throw py::value_error("Cannot sort by both counts and kmers at the same time.");
                 std::unordered_map<uint64_t, uint64_t> counts;
                 for (auto &kv : t.items(0)) counts.emplace(kv.first, kv.second);
                 std::vector<std::pair<std::string, uint64_t>> rows;
                 for (auto &kv : t.hash_to_kmer) {
                     auto it = counts.find(kv.first);
                     if (it != counts.end()) rows.emplace_back(kv.second, it->second);
                 }
                 if (sortkeys) std::sort(rows.begin(), rows.end(), [](auto &a, auto &b) { return a.first < b.first; });
                 else if (sortcounts) std::sort(rows.begin(), rows.end(), [](auto &a, auto &b) { return a.second != b.second ? a.second < b.second : a.first < b.first; });
                 if (!file.is_none()) { write_pairs_tsv(file.cast<std::string>(), rows); rows.clear(); }
                 return rows;
             }, py::arg("file") = py::none(), py::arg("sortcounts") = false, py::arg("sortkeys") = false)
        .def("histo", [](const Table &t, bool zero) {
                 t.sync();
                 uint64_t n = 0;
                 ck(oxg_histo(t.h, nullptr, nullptr, 0, &n));
                 std::vector<uint64_t> f(n + 1), c(n + 1);
                 if (n) ck(oxg_histo(t.h, f.data(), c.data(), n, &n));
                 std::vector<std::pair<uint64_t, uint64_t>> out;
                 if (!zero) {
                     for (uint64_t i = 0; i < n; ++i) out.emplace_back(f[i], c[i]);
                 } else {  // dense 0..=max (src/lib.rs:475-480)
                     const uint64_t mx = n ? f[n - 1] : 0;
                     out.reserve(mx + 1);
                     uint64_t j = 0;
                     for (uint64_t v = 0; v <= mx; ++v) {
                         if (j < n && f[j] == v) out.emplace_back(v, c[j++]);
                         else out.emplace_back(v, 0);
                         if (v == ~0ULL) break;
                     }
                 }
                 return out;
             }, py::arg("zero") = true)
        .def_property_readonly("min", [](const Table &t) { return t.stats().min; })
        .def_property_readonly("max", [](const Table &t) { return t.stats().max; })
        .def_property_readonly("hashes", [](const Table &t) {
                 std::vector<uint64_t> k;
                 for (auto &kv : t.items(0)) k.push_back(kv.first);
                 return k;
             })
        .def_property_readonly("version", [](const Table &t) { return t.version; })
        .def_property_readonly("consumed", [](const Table &t) { return t.consumed; })
        .def_property_readonly("sum_counts", [](const Table &t) { return t.stats().sum; })
        .def_property_readonly("ksize", [](const Table &t) { return (unsigned)t.ksize; })
        .def_property_readonly("store_kmers", [](const Table &t) { return t.store_kmers; })
        .def("consume", &Table::consume, py::arg("seq"), py::arg("skip_bad_kmers") = true)
        .def("consume_many", [](Table &t, const std::vector<std::string> &seqs, bool skip_bad) {
                 std::vector<uint64_t> offs(seqs.size() + 1, 0);
                 for (size_t i = 0; i < seqs.size(); ++i) offs[i + 1] = offs[i] + seqs[i].size();
                 std::string flat;
                 flat.reserve(offs.back());
                 for (auto &s : seqs) flat += s;
                 return t.consume_csr(reinterpret_cast<const uint8_t *>(flat.data()), offs.data(), seqs.size(), skip_bad);
             }, py::arg("seqs"), py::arg("skip_bad_kmers") = true)
        .def("consume_buffer", [](Table &t, py::buffer bases, py::buffer offsets, bool skip_bad) {
                 py::buffer_info b = bases.request(), o = offsets.request();
                 if (b.itemsize != 1 || b.ndim != 1) throw py::value_error("bases must be a 1-D byte buffer");
                 if (o.itemsize != 8 || o.ndim != 1 || o.shape[0] < 1) throw py::value_error("offsets must be a 1-D uint64 buffer with n_reads+1 entries");
                 const uint64_t *op = static_cast<const uint64_t *>(o.ptr);
                 const uint64_t n_reads = (uint64_t)o.shape[0] - 1;
                 if (op[n_reads] > (uint64_t)b.shape[0]) throw py::value_error("offsets run past the end of bases");
                 return t.consume_csr(static_cast<const uint8_t *>(b.ptr), op, n_reads, skip_bad);
             }, py::arg("bases"), py::arg("offsets"), py::arg("skip_bad_kmers") = true)
        .def("consume_file", [](Table &t, const std::string &path, bool skip_bad, uint64_t batch_bytes) {
                 return consume_file(t, path, skip_bad, batch_bytes);
             }, py::arg("path"), py::arg("skip_bad_kmers") = true, py::arg("batch_bytes") = 256ull << 20,
             "Count every record of a FASTA/FASTQ file (plain or gzip); returns (records, k-mers counted)")
        .def("union", [](const Table &a, const Table &b) { return to_pyset(a.setop(b, OXG_UNION)); })
        .def("intersection", [](const Table &a, const Table &b) { return to_pyset(a.setop(b, OXG_INTERSECTION)); })
        .def("difference", [](const Table &a, const Table &b) { return to_pyset(a.setop(b, OXG_DIFFERENCE)); })
        .def("symmetric_difference", [](const Table &a, const Table &b) { return to_pyset(a.setop(b, OXG_SYMMETRIC_DIFFERENCE)); })
        .def("__or__", [](const Table &a, const Table &b) { return to_pyset(a.setop(b, OXG_UNION)); })
        .def("__and__", [](const Table &a, const Table &b) { return to_pyset(a.setop(b, OXG_INTERSECTION)); })
        .def("__sub__", [](const Table &a, const Table &b) { return to_pyset(a.setop(b, OXG_DIFFERENCE)); })
        .def("__xor__", [](const Table &a, const Table &b) { return to_pyset(a.setop(b, OXG_SYMMETRIC_DIFFERENCE)); })
        .def("__iter__", [](const Table &t) {  // iterates a snapshot, like the reference's cloned map (src/lib.rs:658-662)
                 py::list l;
                 for (auto &kv : t.items(0)) l.append(py::make_tuple(kv.first, kv.second));
                 return py::iter(l);
             })
        .def("__len__", &Table::len)
        .def("__getitem__", &Table::get)
        .def("__setitem__", [](Table &t, const std::string &kmer, uint64_t count) {
                 const uint64_t hv = t.hash_kmer(kmer);
                 ck(oxg_set_hash(t.h, hv, count));
             })
        .def("kmers_and_hashes", &Table::kmers_and_hashes, py::arg("seq"), py::arg("skip_bad_kmers") = true)
        .def("jaccard", [](const Table &a, const Table &b) { a.sync(); b.sync(); double d = 0; ck(oxg_jaccard(a.h, b.h, &d)); return d; })
        .def("cosine", [](const Table &a, const Table &b) { a.sync(); b.sync(); double d = 0; ck(oxg_cosine(a.h, b.h, &d)); return d; })
        .def("add", [](Table &a, const Table &b) {  // src/lib.rs:778-837
                 if (a.ksize != b.ksize) throw py::value_error("KmerCountTables must have the same ksize");
                 if (&a == &b) throw std::runtime_error("Already borrowed");
                 uint64_t added = 0, fresh = 0;
                 a.sync(); b.sync();
                 ck(oxg_merge(a.h, b.h, &added, &fresh));
                 a.consumed += b.consumed;
                 if (a.store_kmers) {
                     if (b.store_kmers) for (auto &kv : b.hash_to_kmer) a.hash_to_kmer.emplace(kv.first, kv.second);
                     else fprintf(stderr, "Warning: Incoming table does not store k-mers, but target table does. K-mer information for new hashes will be missing.\n");
                 }
                 printf("Added %llu k-mer counts to the table\n", (unsigned long long)added);
                 printf("Added %llu new keys to the table\n", (unsigned long long)fresh);
                 fflush(stdout);
                 return py::make_tuple(added, fresh);
             }, py::arg("other"))
        .def("reserve", [](Table &t, uint64_t n) { t.sync(); ck(oxg_table_reserve(t.h, n)); }, py::arg("n_keys"))
        .def_property_readonly("device", [](const Table &t) { return t.device; });
}
