// libgci_io.so — native (C++17, zlib, std::thread) decoders into the columnar record schema:
// BAM/BGZF (parallel inflate + parallel record parse, CG:B,I long CIGARs, typed NM), PAF, FASTA N-runs.
// Replaces what the reference gets from pysam / htslib / Biopython (GCI.py:150-166, :201-208, :218-229,
// :28-35), which are not available in the image.  Host code only; the C ABI is include/gci_io.h.
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/gci_io.h"

namespace {

thread_local std::string g_err;

int fail(const std::string& m) {
  g_err = m;
  return -1;
}

struct MappedFile {
  const uint8_t* p = nullptr;
  size_t n = 0;
  int fd = -1;
  bool open(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) return false;
    n = (size_t)st.st_size;
    if (n == 0) { p = nullptr; return true; }
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) return false;
    p = (const uint8_t*)m;
    madvise(m, n, MADV_SEQUENTIAL);
    return true;
  }
  ~MappedFile() {
    if (p) munmap((void*)p, n);
    if (fd >= 0) ::close(fd);
  }
};

template <typename F>
void parallel_for(int64_t n, int threads, F&& fn) {
  threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n));
  if (threads == 1) { fn(0, n, 0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) th.emplace_back([&, t] { fn(n * t / threads, n * (t + 1) / threads, t); });
  for (auto& x : th) x.join();
}

inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }

}  // namespace

// Read name -> dense id in order of first appearance, shared by all files of a read type.  Open addressing over
// (hash, arena offset): a lookup touches one cache line and allocates nothing; the 64-bit hashes are computed by
// the decoders' parallel phases, only the probing runs in file order.
inline uint64_t name_hash(const char* s, size_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, s, 8);
    h = (h ^ w) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    s += 8;
    n -= 8;
  }
  uint64_t w = 0;
  memcpy(&w, s, n);
  h = (h ^ w) * 0xD6E8FEB86659FD93ull;
  h ^= h >> 29;
  h *= 0x94D049BB133111EBull;
  return h ^ (h >> 32);
}

struct gci_interner {
  struct Slot { uint64_t hash, off; uint32_t len, id; };   // id == UINT32_MAX: empty
  std::vector<Slot> slots;
  std::vector<char> arena;                  // the distinct names back to back
  uint32_t count = 0;

  void grow() {
    const size_t cap = slots.empty() ? (size_t)1 << 16 : slots.size() * 2;
    std::vector<Slot> fresh(cap, Slot{0, 0, 0, UINT32_MAX});
    for (const Slot& s : slots) {
      if (s.id == UINT32_MAX) continue;
      size_t i = (size_t)s.hash & (cap - 1);
      while (fresh[i].id != UINT32_MAX) i = (i + 1) & (cap - 1);
      fresh[i] = s;
    }
    slots.swap(fresh);
  }
  uint32_t intern(const char* s, size_t n, uint64_t h) {
    if ((size_t)(count + 1) * 10 > slots.size() * 7) grow();          // load factor <= 0.7
    const size_t mask = slots.size() - 1;
    for (size_t i = (size_t)h & mask;; i = (i + 1) & mask) {
      Slot& sl = slots[i];
      if (sl.id == UINT32_MAX) {
        sl = Slot{h, (uint64_t)arena.size(), (uint32_t)n, count};
        arena.insert(arena.end(), s, s + n);
        return count++;
      }
      if (sl.hash == h && sl.len == n && memcmp(arena.data() + sl.off, s, n) == 0) return sl.id;
    }
  }
};

// One BAM file decoded into columns.  The file is never inflated as a whole: gci_bam_open streams it through a
// window of a few hundred MB (parallel inflate of the window's BGZF blocks, sequential walk over the record
// sizes, parallel field extraction into the columns; a record cut by the window border is carried into the
// next window), so host memory stays bounded for BAMs of any size.
struct gci_bam {
  std::vector<std::string> ref_names;
  std::vector<int64_t> ref_lens;
  std::vector<int32_t> ref_id, ref_start, nm, qlen;
  std::vector<uint8_t> mapq;
  std::vector<uint16_t> flag;
  std::vector<uint64_t> cig_off;             // [n+1]
  std::vector<uint32_t> cigar;
  std::vector<uint64_t> name_off;            // [n+1] into names
  std::vector<char> names;                   // read names back to back (no terminators)
  std::vector<uint64_t> name_hash;           // [n] name_hash() of every read name
  int threads = 1;
};

namespace {
// inflated bytes per window (plus the carried tail); GCI_IO_WINDOW_BYTES overrides it (tests use tiny windows)
inline size_t bam_window_bytes() {
  const char* e = getenv("GCI_IO_WINDOW_BYTES");
  const long long v = e ? atoll(e) : 0;
  return v > 0 ? (size_t)v : size_t(256) << 20;
}

// number of CIGAR ops of the record at r (CG:B,I long-CIGAR convention, SAM spec §4.2.2); *cg = first op of the
// CG payload or NULL
inline uint64_t bam_record_ops(const uint8_t* r, const uint8_t** cg) {
  *cg = nullptr;
  const uint32_t n_cig = rd16(r + 16);
  if (n_cig != 2) return n_cig;
  const int32_t l_seq = rdi32(r + 20);
  const uint8_t l_name = r[12];
  const uint8_t* cig = r + 36 + l_name;
  const uint32_t o0 = rd32(cig), o1 = rd32(cig + 4);
  if (!((o0 & 15) == 4 && (int32_t)(o0 >> 4) == l_seq && (o1 & 15) == 3)) return n_cig;
  const uint8_t* end = r + 4 + rdi32(r);
  const uint8_t* x = cig + 8 + (l_seq + 1) / 2 + l_seq;
  while (x + 3 <= end) {
    const uint8_t t = x[2];
    if (x[0] == 'C' && x[1] == 'G' && t == 'B' && x[3] == 'I') {
      *cg = x + 8;
      return rd32(x + 4);
    }
    x += 3;
    if (t == 'A' || t == 'c' || t == 'C') x += 1;
    else if (t == 's' || t == 'S') x += 2;
    else if (t == 'i' || t == 'I' || t == 'f') x += 4;
    else if (t == 'Z' || t == 'H') { while (x < end && *x) x++; x++; }
    else if (t == 'B') {
      const uint8_t st = x[0];
      const uint32_t cnt = rd32(x + 1);
      const int sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
      x += 5 + (size_t)cnt * sz;
    } else break;
  }
  return n_cig;
}

// NM aux value of the record at r (any integer type), INT32_MIN if absent
inline int32_t bam_record_nm(const uint8_t* r) {
  const uint8_t* end = r + 4 + rdi32(r);
  const uint8_t l_name = r[12];
  const uint32_t n_cig = rd16(r + 16);
  const int32_t l_seq = rdi32(r + 20);
  int32_t nmv = INT32_MIN;
  const uint8_t* x = r + 36 + l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq;
  while (x + 3 <= end) {
    const uint8_t t = x[2];
    const bool is_nm = x[0] == 'N' && x[1] == 'M';
    x += 3;
    if (t == 'A' || t == 'c' || t == 'C') { if (is_nm) nmv = t == 'c' ? (int8_t)x[0] : x[0]; x += 1; }
    else if (t == 's' || t == 'S') { if (is_nm) nmv = t == 's' ? (int16_t)rd16(x) : rd16(x); x += 2; }
    else if (t == 'i' || t == 'I' || t == 'f') { if (is_nm && t != 'f') nmv = rdi32(x); x += 4; }
    else if (t == 'Z' || t == 'H') { while (x < end && *x) x++; x++; }
    else if (t == 'B') {
      const uint8_t st = x[0];
      const uint32_t cnt = rd32(x + 1);
      const int sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
      x += 5 + (size_t)cnt * sz;
    } else break;
    if (is_nm && nmv != INT32_MIN) break;
  }
  return nmv;
}
}  // namespace

extern "C" {

const char* gci_io_last_error(void) { return g_err.c_str(); }

gci_interner* gci_interner_create(void) { return new gci_interner(); }
void gci_interner_destroy(gci_interner* it) { delete it; }
int64_t gci_interner_size(gci_interner* it) { return it ? (int64_t)it->count : 0; }

// ---- BAM ------------------------------------------------------------------------------------------
int gci_bam_open(const char* path, int threads, gci_bam** out) {
  if (!path || !out) return fail("bad argument");
  *out = nullptr;
  MappedFile f;
  if (!f.open(path)) return fail(std::string("cannot open ") + path);
  // pass 1: walk the BGZF block headers (BSIZE in the BC extra field, ISIZE in the trailer)
  struct Blk { size_t src, clen, ulen; };
  std::vector<Blk> blocks;
  size_t pos = 0;
  while (pos + 18 <= f.n) {
    const uint8_t* b = f.p + pos;
    if (!(b[0] == 0x1f && b[1] == 0x8b && b[2] == 8 && (b[3] & 4))) return fail("not a BGZF block");
    const size_t xlen = rd16(b + 10);
    size_t bsize = 0;
    for (size_t k = 0; k + 4 <= xlen;) {
      const uint8_t* e = b + 12 + k;
      const size_t slen = rd16(e + 2);
      if (e[0] == 66 && e[1] == 67 && slen == 2) bsize = (size_t)rd16(e + 4) + 1;
      k += 4 + slen;
    }
    if (bsize == 0 || pos + bsize > f.n) return fail("BGZF block without BSIZE / truncated file");
    blocks.push_back({pos + 12 + xlen, bsize - 12 - xlen - 8, (size_t)rd32(b + bsize - 4)});
    pos += bsize;
  }
  std::unique_ptr<gci_bam> bam(new gci_bam());
  bam->threads = std::max(1, threads);
  bam->cig_off.push_back(0);
  bam->name_off.push_back(0);

  // pass 2: window by window
  uint8_t* buf = nullptr;                    // [carried tail of the previous window | this window, inflated]
  size_t buf_cap = 0, carry = 0;
  struct Free { uint8_t*& p; ~Free() { free(p); } } free_buf{buf};
  bool header_done = false;
  const size_t window_bytes = bam_window_bytes();
  std::vector<size_t> dst, rec_off;
  std::vector<uint64_t> ops;
  std::vector<const uint8_t*> cg;
  for (size_t w0 = 0; w0 < blocks.size();) {
    size_t w1 = w0, win = 0;
    dst.clear();
    while (w1 < blocks.size() && (w1 == w0 || win + blocks[w1].ulen <= window_bytes)) {
      dst.push_back(win);
      win += blocks[w1].ulen;
      w1++;
    }
    if (carry + win > buf_cap) {
      const size_t want = carry + win + (carry + win) / 8;
      uint8_t* nb = (uint8_t*)malloc(want);       // malloc + memcpy of the tail: realloc would copy everything
      if (!nb) return fail("out of memory");
      if (carry) memcpy(nb, buf, carry);
      free(buf);
      buf = nb;
      buf_cap = want;
    }
    std::atomic<int> bad{0};
    parallel_for((int64_t)(w1 - w0), bam->threads, [&](int64_t a, int64_t b, int) {
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
      for (int64_t i = a; i < b; i++) {
        const Blk& k = blocks[w0 + i];
        if (k.ulen == 0) continue;
        inflateReset(&zs);
        zs.next_in = const_cast<Bytef*>(f.p + k.src);
        zs.avail_in = (uInt)k.clen;
        zs.next_out = buf + carry + dst[i];
        zs.avail_out = (uInt)k.ulen;
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0) { bad = 1; break; }
      }
      inflateEnd(&zs);
    });
    if (bad) return fail("BGZF inflate failed");
    const size_t avail = carry + win;
    size_t p = 0;
    if (!header_done) {
      // magic, l_text, text, n_ref, (l_name, name, l_ref) x n_ref: parse only once it is completely here
      bool complete = false;
      do {
        if (avail < 12) break;
        if (memcmp(buf, "BAM\1", 4) != 0) return fail("not a BAM file");
        const int32_t l_text = rdi32(buf + 4);
        if (l_text < 0) return fail("corrupt BAM header");
        size_t q = 8 + (size_t)l_text;
        if (q + 4 > avail) break;
        const int32_t n_ref = rdi32(buf + q);
        if (n_ref < 0) return fail("corrupt BAM header");
        q += 4;
        std::vector<std::string> names;
        std::vector<int64_t> lens;
        bool cut = false;
        for (int32_t i = 0; i < n_ref; i++) {
          if (q + 4 > avail) { cut = true; break; }
          const int32_t l_name = rdi32(buf + q);
          if (l_name < 0) return fail("corrupt BAM header");
          if (q + 8 + (size_t)l_name > avail) { cut = true; break; }
          names.emplace_back((const char*)buf + q + 4, (size_t)std::max(0, l_name - 1));
          lens.push_back(rdi32(buf + q + 4 + l_name));
          q += 8 + (size_t)l_name;
        }
        if (cut) break;
        bam->ref_names.swap(names);
        bam->ref_lens.swap(lens);
        p = q;
        complete = true;
      } while (false);
      if (!complete) {                           // header longer than what has been inflated so far
        carry = avail;
        w0 = w1;
        continue;
      }
      header_done = true;
    }
    // record boundaries (sequential: each record names its own size)
    rec_off.clear();
    while (p + 4 <= avail) {
      const int32_t bs = rdi32(buf + p);
      if (bs < 32) return fail("corrupt BAM record");
      if (p + 4 + (size_t)bs > avail) break;      // cut by the window border: carried over
      rec_off.push_back(p);
      p += 4 + (size_t)bs;
    }
    const int64_t m = (int64_t)rec_off.size();
    const size_t n0 = bam->ref_id.size();
    // CIGAR and name sizes of the window's records, then their places in the columns
    ops.assign((size_t)m, 0);
    cg.assign((size_t)m, nullptr);
    parallel_for(m, bam->threads, [&](int64_t a, int64_t b, int) {
      for (int64_t i = a; i < b; i++) ops[i] = bam_record_ops(buf + rec_off[i], &cg[i]);
    });
    bam->cig_off.resize(n0 + m + 1);
    bam->name_off.resize(n0 + m + 1);
    for (int64_t i = 0; i < m; i++) {
      const uint8_t l_name = buf[rec_off[i] + 12];
      bam->cig_off[n0 + i + 1] = bam->cig_off[n0 + i] + ops[i];
      bam->name_off[n0 + i + 1] = bam->name_off[n0 + i] + (l_name ? (uint64_t)l_name - 1 : 0);
    }
    bam->ref_id.resize(n0 + m); bam->ref_start.resize(n0 + m); bam->nm.resize(n0 + m); bam->qlen.resize(n0 + m);
    bam->mapq.resize(n0 + m); bam->flag.resize(n0 + m); bam->name_hash.resize(n0 + m);
    bam->cigar.resize((size_t)bam->cig_off[n0 + m]);
    bam->names.resize((size_t)bam->name_off[n0 + m]);
    parallel_for(m, bam->threads, [&](int64_t a, int64_t b, int) {
      for (int64_t i = a; i < b; i++) {
        const uint8_t* r = buf + rec_off[i];
        const size_t k = n0 + (size_t)i;
        const uint8_t l_name = r[12];
        bam->ref_id[k] = rdi32(r + 4);
        bam->ref_start[k] = rdi32(r + 8);
        bam->mapq[k] = r[13];
        bam->flag[k] = rd16(r + 18);
        bam->qlen[k] = rdi32(r + 20);
        const uint8_t* cig = cg[i] ? cg[i] : r + 36 + l_name;
        memcpy(bam->cigar.data() + bam->cig_off[k], cig, 4 * (size_t)ops[i]);
        bam->nm[k] = bam_record_nm(r);
        const size_t name_n = (size_t)(bam->name_off[k + 1] - bam->name_off[k]);
        memcpy(bam->names.data() + bam->name_off[k], r + 36, name_n);
        bam->name_hash[k] = name_hash((const char*)r + 36, name_n);
      }
    });
    carry = avail - p;
    if (carry) memmove(buf, buf + p, carry);
    w0 = w1;
  }
  if (!header_done) return fail(blocks.empty() ? "not a BAM file" : "truncated BAM header");
  if (carry) return fail("corrupt BAM record");   // bytes after the last complete record
  *out = bam.release();
  return 0;
}

void gci_bam_close(gci_bam* b) { delete b; }
int32_t gci_bam_n_refs(gci_bam* b) { return b ? (int32_t)b->ref_names.size() : 0; }
const char* gci_bam_ref_name(gci_bam* b, int32_t i) { return b->ref_names[i].c_str(); }
int64_t gci_bam_ref_len(gci_bam* b, int32_t i) { return b->ref_lens[i]; }
int64_t gci_bam_n_records(gci_bam* b) { return b ? (int64_t)b->ref_id.size() : 0; }
int64_t gci_bam_n_ops(gci_bam* b) { return b ? (int64_t)b->cigar.size() : 0; }

int gci_bam_fill(gci_bam* b, gci_interner* it, int32_t* ref_id, int32_t* ref_start, uint8_t* mapq, uint16_t* flag,
                 int32_t* nm, int32_t* qlen, uint32_t* read_id, uint64_t* cigar_off, uint32_t* cigar) {
  if (!b || !it) return fail("bad argument");
  const size_t n = b->ref_id.size();
  memcpy(cigar_off, b->cig_off.data(), sizeof(uint64_t) * (n + 1));
  if (n) {
    memcpy(ref_id, b->ref_id.data(), 4 * n);
    memcpy(ref_start, b->ref_start.data(), 4 * n);
    memcpy(mapq, b->mapq.data(), n);
    memcpy(flag, b->flag.data(), 2 * n);
    memcpy(nm, b->nm.data(), 4 * n);
    memcpy(qlen, b->qlen.data(), 4 * n);
  }
  if (!b->cigar.empty()) memcpy(cigar, b->cigar.data(), 4 * b->cigar.size());
  // read names -> dense ids, in file order (sequential: one shared table per read type)
  for (size_t i = 0; i < n; i++)
    read_id[i] = it->intern(b->names.data() + b->name_off[i], (size_t)(b->name_off[i + 1] - b->name_off[i]),
                            b->name_hash[i]);
  return 0;
}

// ---- PAF ------------------------------------------------------------------------------------------
struct gci_paf {
  std::vector<uint32_t> read_id;
  std::vector<int32_t> cols[9];   // qlen qstart qend ref_id tstart tend nmatch alnlen mapq
};

static inline bool parse_int(std::string_view s, long long& v) {
  // Python int(): surrounding blanks are accepted (a tab-split field may keep them), optional sign, '_' between digits
  while (!s.empty() && (s.front() == ' ' || s.front() == '\r' || s.front() == '\f' || s.front() == '\v')) s.remove_prefix(1);
  while (!s.empty() && (s.back() == ' ' || s.back() == '\r' || s.back() == '\f' || s.back() == '\v')) s.remove_suffix(1);
  if (s.empty()) return false;
  size_t i = 0;
  bool neg = false;
  if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
  if (i >= s.size()) return false;
  long long x = 0;
  for (; i < s.size(); i++) {
    if (s[i] == '_') continue;
    if (s[i] < '0' || s[i] > '9') return false;
    x = x * 10 + (s[i] - '0');
    if (x > (1ll << 40)) return false;            // far outside any column: no overflow of x itself
  }
  v = neg ? -x : x;
  return true;
}

// host threads for the text decoders (PAF lines, FASTA); GCI_IO_THREADS overrides the hardware count
static int io_threads() {
  const char* e = getenv("GCI_IO_THREADS");
  int t = e ? atoi(e) : 0;
  if (t <= 0) t = (int)std::thread::hardware_concurrency();
  return std::max(1, std::min(t, 64));
}

// The file is cut into one chunk per thread at line starts; every thread parses its lines into its own
// columns (read names as views into the mapping), the parts are appended in file order and the names are
// interned sequentially, so ids follow first appearance exactly as in a single pass.
int gci_paf_open(const char* path, gci_interner* it, int32_t n_contigs, const char* const* contig_names,
                 gci_paf** out) {
  if (!path || !it || !out) return fail("bad argument");
  *out = nullptr;
  MappedFile f;
  if (!f.open(path)) return fail(std::string("cannot open ") + path);
  std::unordered_map<std::string_view, int32_t> cidx;
  for (int32_t i = 0; i < n_contigs; i++) cidx.emplace(std::string_view(contig_names[i]), i);
  const char* const begin = (const char*)f.p;
  const char* const end = begin + f.n;
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(io_threads(), (int64_t)(f.n >> 16) + 1));
  std::vector<const char*> cut((size_t)T + 1, end);
  cut[0] = begin;
  for (int t = 1; t < T; t++) {
    const char* q = begin + f.n / (size_t)T * (size_t)t;
    if (q < cut[t - 1]) q = cut[t - 1];
    const char* nl = q < end ? (const char*)memchr(q, '\n', (size_t)(end - q)) : nullptr;
    cut[t] = nl ? nl + 1 : end;
  }
  struct Part {
    std::vector<int32_t> cols[9];   // qlen qstart qend ref_id tstart tend nmatch alnlen mapq
    std::vector<std::string_view> names;
    std::vector<uint64_t> hashes;
    int64_t lines = 0, bad_line = -1;
    const char* bad_what = nullptr;
  };
  std::vector<Part> parts((size_t)T);
  parallel_for(T, T, [&](int64_t t0, int64_t t1, int) {
    for (int64_t t = t0; t < t1; t++) {
      Part& pt = parts[(size_t)t];
      const char* p = cut[(size_t)t];
      const char* const pe = cut[(size_t)t + 1];
      while (p < pe) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(pe - p));
        const char* le = nl ? nl : pe;
        pt.lines++;
        // line.strip().split("\t")
        const char* a = p;
        const char* b = le;
        while (a < b && (*a == ' ' || *a == '\t' || *a == '\r' || *a == '\v' || *a == '\f')) a++;
        while (b > a && (b[-1] == ' ' || b[-1] == '\t' || b[-1] == '\r' || b[-1] == '\v' || b[-1] == '\f')) b--;
        std::string_view col[12];
        int nc = 0;
        const char* s = a;
        for (const char* q = a; q <= b && nc < 12; q++) {
          if (q == b || *q == '\t') {
            col[nc++] = std::string_view(s, (size_t)(q - s));
            s = q + 1;
          }
        }
        if (nc < 12) { pt.bad_line = pt.lines; pt.bad_what = "fewer than 12 columns"; break; }
        long long v[12];
        static const int want[9] = {1, 2, 3, 7, 8, 9, 10, 11, -1};
        bool ok = true;
        for (int k = 0; want[k] >= 0; k++) ok = ok && parse_int(col[want[k]], v[want[k]]);
        if (!ok) { pt.bad_line = pt.lines; pt.bad_what = "invalid integer"; break; }
        // the device columns are int32: a value outside that range is reported, not truncated
        for (int k = 0; want[k] >= 0; k++) ok = ok && v[want[k]] >= INT32_MIN && v[want[k]] <= INT32_MAX;
        if (!ok) { pt.bad_line = pt.lines; pt.bad_what = "integer outside the int32 range of the device columns"; break; }
        pt.names.push_back(col[0]);
        pt.hashes.push_back(name_hash(col[0].data(), col[0].size()));
        auto c = cidx.find(col[5]);
        pt.cols[0].push_back((int32_t)v[1]);
        pt.cols[1].push_back((int32_t)v[2]);
        pt.cols[2].push_back((int32_t)v[3]);
        pt.cols[3].push_back(c == cidx.end() ? -1 : c->second);
        pt.cols[4].push_back((int32_t)v[7]);
        pt.cols[5].push_back((int32_t)v[8]);
        pt.cols[6].push_back((int32_t)v[9]);
        pt.cols[7].push_back((int32_t)v[10]);
        pt.cols[8].push_back((int32_t)v[11]);
        p = nl ? nl + 1 : pe;
      }
    }
  });
  int64_t line0 = 0;
  size_t total = 0;
  for (const Part& pt : parts) {                // the first offending line in file order, like a single pass
    if (pt.bad_line >= 0) return fail("PAF line " + std::to_string(line0 + pt.bad_line) + ": " + pt.bad_what);
    line0 += pt.lines;
    total += pt.names.size();
  }
  std::unique_ptr<gci_paf> paf(new gci_paf());
  paf->read_id.reserve(total);
  for (auto& c : paf->cols) c.reserve(total);
  for (const Part& pt : parts) {
    for (int k = 0; k < 9; k++) paf->cols[k].insert(paf->cols[k].end(), pt.cols[k].begin(), pt.cols[k].end());
    for (size_t i = 0; i < pt.names.size(); i++)
      paf->read_id.push_back(it->intern(pt.names[i].data(), pt.names[i].size(), pt.hashes[i]));
  }
  *out = paf.release();
  return 0;
}

int64_t gci_paf_n_lines(gci_paf* p) { return p ? (int64_t)p->read_id.size() : 0; }
void gci_paf_close(gci_paf* p) { delete p; }
int gci_paf_fill(gci_paf* p, uint32_t* read_id, int32_t* qlen, int32_t* qstart, int32_t* qend, int32_t* ref_id,
                 int32_t* tstart, int32_t* tend, int32_t* nmatch, int32_t* alnlen, int32_t* mapq) {
  if (!p) return fail("bad argument");
  const size_t n = p->read_id.size();
  memcpy(read_id, p->read_id.data(), 4 * n);
  int32_t* dst[9] = {qlen, qstart, qend, ref_id, tstart, tend, nmatch, alnlen, mapq};
  for (int k = 0; k < 9; k++) memcpy(dst[k], p->cols[k].data(), 4 * n);
  return 0;
}

// ---- FASTA: record ids and N/n runs (GCI.py:28-35, :939-941) ------------------------------------------
struct gci_fasta {
  std::vector<std::string> ids;
  std::vector<int32_t> run_rec;
  std::vector<int64_t> run_start, run_end;
};

// Line by line: a sequence line without N / n and without blanks (nearly all of them) only moves the position
// counter -- five memchr passes instead of a byte loop -- and only a line that holds an N or a blank is walked
// byte by byte.  gzread serves plain and gzip files alike; an incomplete last line of a buffer is carried over.
int gci_fasta_open(const char* path, gci_fasta** out) {
  if (!path || !out) return fail("bad argument");
  *out = nullptr;
  gzFile g = gzopen(path, "rb");
  if (!g) return fail(std::string("cannot open ") + path);
  gzbuffer(g, 1 << 20);
  std::unique_ptr<gci_fasta> fa(new gci_fasta());
  int64_t pos = 0, run0 = -1;
  auto close_run = [&]() {
    if (run0 >= 0) {
      fa->run_rec.push_back((int32_t)fa->ids.size() - 1);
      fa->run_start.push_back(run0);
      fa->run_end.push_back(pos);
      run0 = -1;
    }
  };
  auto line = [&](const char* s, const char* e) {          // one line without its '\n'
    if (s < e && *s == '>') {
      close_run();
      const char* a = s + 1;
      while (a < e && (*a == ' ' || *a == '\t')) a++;
      const char* b = a;
      while (b < e && *b != ' ' && *b != '\t' && *b != '\r') b++;
      fa->ids.emplace_back(a, (size_t)(b - a));
      pos = 0;
      return;
    }
    if (fa->ids.empty() || s == e) return;                   // text before the first header / empty line
    if (e[-1] == '\r') e--;                                  // CRLF files: the common blank, dealt with once
    const size_t n = (size_t)(e - s);
    if (n == 0) return;
    if (!memchr(s, 'N', n) && !memchr(s, 'n', n) && !memchr(s, ' ', n) && !memchr(s, '\t', n) && !memchr(s, '\r', n)) {
      close_run();
      pos += (int64_t)n;
      return;
    }
    for (const char* p = s; p < e; p++) {
      const char c = *p;
      if (c == '\r' || c == ' ' || c == '\t') continue;      // line.strip()
      if (c == 'N' || c == 'n') { if (run0 < 0) run0 = pos; }
      else close_run();
      pos++;
    }
  };
  std::vector<char> buf((size_t)1 << 22);
  size_t have = 0;                                           // bytes of an incomplete line at the front of buf
  for (;;) {
    if (have == buf.size()) buf.resize(buf.size() * 2);      // a line longer than the buffer (unwrapped FASTA)
    const int got = gzread(g, buf.data() + have, (unsigned)std::min<size_t>(buf.size() - have, (size_t)1 << 30));
    if (got < 0) { gzclose(g); return fail("read error"); }
    if (got == 0) break;
    const char* p = buf.data();
    const char* const end = p + have + (size_t)got;
    const char* scan = p + have;                             // the carried part holds no newline
    for (;;) {
      const char* nl = (const char*)memchr(scan, '\n', (size_t)(end - scan));
      if (!nl) break;
      line(p, nl);
      p = nl + 1;
      scan = p;
    }
    have = (size_t)(end - p);
    if (have && p != buf.data()) memmove(buf.data(), p, have);
  }
  if (have) line(buf.data(), buf.data() + have);             // last line without a newline
  close_run();
  gzclose(g);
  *out = fa.release();
  return 0;
}

int32_t gci_fasta_n_records(gci_fasta* f) { return f ? (int32_t)f->ids.size() : 0; }
const char* gci_fasta_id(gci_fasta* f, int32_t i) { return f->ids[i].c_str(); }
int64_t gci_fasta_n_runs(gci_fasta* f) { return f ? (int64_t)f->run_rec.size() : 0; }
int gci_fasta_runs(gci_fasta* f, int32_t* rec, int64_t* start, int64_t* end) {
  if (!f) return fail("bad argument");
  const size_t n = f->run_rec.size();
  memcpy(rec, f->run_rec.data(), 4 * n);
  memcpy(start, f->run_start.data(), 8 * n);
  memcpy(end, f->run_end.data(), 8 * n);
  return 0;
}
void gci_fasta_close(gci_fasta* f) { delete f; }

// ---- .depth.gz: ">name" lines followed by one decimal depth per line (GCI.py:110-117; reader contract
// utility/GCI_score.py:25-37) ------------------------------------------------------------------------------
// The (multi-member) gzip stream is inflated through a window; the number tokens of a window are counted and
// then parsed by all threads, each writing its slice of the contig's array.
// growable int32 array without the zero fill of std::vector::resize (realloc of a large block is a remap)
struct DepthArray {
  int32_t* p = nullptr;
  size_t n = 0, cap = 0;
  DepthArray() = default;
  DepthArray(const DepthArray&) = delete;
  DepthArray& operator=(const DepthArray&) = delete;
  DepthArray(DepthArray&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
  ~DepthArray() { free(p); }
  bool grow(size_t extra) {
    if (n + extra > cap) {
      const size_t want = std::max(n + extra, cap + cap / 2 + 1024);
      int32_t* q = (int32_t*)realloc(p, want * sizeof(int32_t));
      if (!q) return false;
      p = q;
      cap = want;
    }
    n += extra;
    return true;
  }
};

struct gci_depth {
  std::vector<std::string> names;
  std::vector<DepthArray> depth;
};

namespace {
// blank = anything up to ' ' (newline, CR, tab, VT, FF, space; other control bytes are not valid in the file anyway)
inline bool is_blank(char c) { return (unsigned char)c <= ' '; }

// tokens (maximal runs of non-blank bytes) that START inside [a, b); branch-free so that it vectorises
int64_t count_tokens(const char* base, const char* a, const char* b) {
  if (a >= b) return 0;
  int64_t n = (a == base || is_blank(a[-1])) && !is_blank(a[0]);
  const unsigned char* u = (const unsigned char*)a;
  const int64_t len = b - a;
  for (int64_t i = 1; i < len; i++) n += (u[i] > ' ') & (u[i - 1] <= ' ');
  return n;
}

// parse the tokens that start inside [a, b) into out[0..]; returns false on a token that is not a Python int
bool parse_tokens(const char* base, const char* a, const char* b, const char* end, int32_t* out) {
  const char* p = a;
  if (p != base && !is_blank(p[-1]))            // in the middle of a token that belongs to the previous slice
    while (p < b && !is_blank(*p)) p++;
  for (;;) {
    while (p < b && is_blank(*p)) p++;
    if (p >= b) break;
    // fast path: plain digits up to a blank (or the end of the data)
    uint32_t x = 0;
    const char* q = p;
    unsigned d;
    while (q < end && (d = (unsigned)(*q - '0')) <= 9u && q - p < 9) { x = x * 10u + d; q++; }
    if (q > p && (q == end || is_blank(*q))) {
      *out++ = (int32_t)x;
      p = q;
      continue;
    }
    // anything else Python's int() takes: sign, underscores, ten digits
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; p++; }
    long long v = 0;
    int digits = 0;
    while (p < end && !is_blank(*p)) {
      if (*p == '_') { p++; continue; }
      if (*p < '0' || *p > '9') return false;
      v = v * 10 + (*p - '0');
      if (v > 0x7fffffffll + 1) return false;     // does not fit the int32 depth tracks
      digits++;
      p++;
    }
    if (!digits) return false;
    v = neg ? -v : v;
    if (v > 0x7fffffffll) return false;
    *out++ = (int32_t)v;
  }
  return true;
}
}  // namespace

int gci_depth_open(const char* path, int threads, gci_depth** out) {
  if (!path || !out) return fail("bad argument");
  *out = nullptr;
  MappedFile f;
  if (!f.open(path)) return fail(std::string("cannot open ") + path);
  threads = std::max(1, threads);
  std::unique_ptr<gci_depth> dp(new gci_depth());
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, 15 + 32) != Z_OK) return fail("zlib init failed");
  struct End { z_stream* z; ~End() { inflateEnd(z); } } end_zs{&zs};
  const size_t window = bam_window_bytes();
  std::vector<char> buf(std::max<size_t>(window, 1 << 16) + (1 << 16));
  size_t have = 0, in_pos = 0;
  bool stream_done = f.n == 0;
  std::atomic<int> bad{0};
  auto numbers = [&](const char* a, const char* b) -> bool {      // [a, b): complete lines without a header
    if (a >= b) return true;
    const int64_t span = b - a;
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(threads, span >> 16));
    std::vector<int64_t> cnt((size_t)T + 1, 0);
    parallel_for(T, T, [&](int64_t t0, int64_t t1, int) {
      for (int64_t t = t0; t < t1; t++) cnt[(size_t)t + 1] = count_tokens(a, a + span * t / T, a + span * (t + 1) / T);
    });
    for (int t = 0; t < T; t++) cnt[(size_t)t + 1] += cnt[(size_t)t];
    if (cnt[(size_t)T] == 0) return true;
    if (dp->depth.empty()) return false;                           // numbers before the first ">name" line
    DepthArray& v = dp->depth.back();
    const size_t n0 = v.n;
    if (!v.grow((size_t)cnt[(size_t)T])) return false;
    parallel_for(T, T, [&](int64_t t0, int64_t t1, int) {
      for (int64_t t = t0; t < t1; t++)
        if (!parse_tokens(a, a + span * t / T, a + span * (t + 1) / T, b, v.p + n0 + cnt[(size_t)t])) bad = 1;
    });
    return bad == 0;
  };
  auto lines = [&](const char* a, const char* b) -> bool {          // [a, b): complete lines
    const char* p = a;
    while (p < b) {
      const char* h = (const char*)memchr(p, '>', (size_t)(b - p));
      if (!h) return numbers(p, b);
      const char* ls = h;                                           // start of the header's line
      while (ls > p && ls[-1] != '\n') ls--;
      if (!numbers(p, ls)) return false;
      const char* nl = (const char*)memchr(h, '\n', (size_t)(b - h));
      const char* le = nl ? nl : b;
      // item = line.strip(); must start with '>'; target = item.split('>')[-1]
      const char* s0 = ls;
      const char* s1 = le;
      while (s0 < s1 && is_blank(*s0)) s0++;
      while (s1 > s0 && is_blank(s1[-1])) s1--;
      if (s0 >= s1 || *s0 != '>') return false;
      const char* last = s1;
      while (last > s0 && last[-1] != '>') last--;
      dp->names.emplace_back(last, (size_t)(s1 - last));
      dp->depth.emplace_back();
      p = nl ? nl + 1 : b;
    }
    return true;
  };
  while (!stream_done || have) {
    // fill the window
    while (!stream_done && have < window) {
      zs.next_in = const_cast<Bytef*>(f.p + in_pos);
      zs.avail_in = (uInt)std::min<size_t>(f.n - in_pos, (size_t)1 << 30);
      zs.next_out = (Bytef*)buf.data() + have;
      zs.avail_out = (uInt)std::min<size_t>(buf.size() - have, (size_t)1 << 30);
      const uInt in0 = zs.avail_in, out0 = zs.avail_out;
      const int rc = inflate(&zs, Z_NO_FLUSH);
      in_pos += in0 - zs.avail_in;
      have += out0 - zs.avail_out;
      if (rc == Z_STREAM_END) {
        if (in_pos >= f.n) stream_done = true;
        else if (inflateReset(&zs) != Z_OK) return fail("zlib reset failed");   // next gzip member
      } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
        return fail("not a gzip stream / corrupt .depth.gz");
      } else if (in0 == zs.avail_in && out0 == zs.avail_out) {
        if (in_pos >= f.n) return fail("truncated .depth.gz");
        break;                                                       // output window full
      }
      if (have >= buf.size()) break;
    }
    // complete lines of the window; the rest is carried over (everything, once the stream has ended)
    size_t upto = have;
    if (!stream_done) {
      while (upto > 0 && buf[upto - 1] != '\n') upto--;
      if (upto == 0) {                                               // one line longer than the window
        if (have == buf.size()) buf.resize(buf.size() * 2);
        continue;
      }
    }
    if (!lines(buf.data(), buf.data() + upto)) return fail("malformed .depth.gz (a line is neither '>name' nor an integer)");
    have -= upto;
    if (have) memmove(buf.data(), buf.data() + upto, have);
    if (stream_done && have == 0) break;
  }
  *out = dp.release();
  return 0;
}

int32_t gci_depth_n_contigs(gci_depth* d) { return d ? (int32_t)d->names.size() : 0; }
const char* gci_depth_name(gci_depth* d, int32_t i) { return d->names[(size_t)i].c_str(); }
int64_t gci_depth_len(gci_depth* d, int32_t i) { return (int64_t)d->depth[(size_t)i].n; }
int gci_depth_fill(gci_depth* d, int32_t i, int32_t* out) {
  if (!d || i < 0 || (size_t)i >= d->depth.size()) return fail("bad argument");
  const DepthArray& v = d->depth[(size_t)i];
  if (v.n) memcpy(out, v.p, 4 * v.n);
  return 0;
}
void gci_depth_close(gci_depth* d) { delete d; }

}  // extern "C"
