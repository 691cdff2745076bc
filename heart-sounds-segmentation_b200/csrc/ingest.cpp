// Recording ingest: batch parser for the dataset's on-disk format.
//
// Replaces the per-file pandas.read_csv of reference hss/datasets/heart_sounds.py:193-197 (`_load_file`): a two-column CSV whose
// first row is a header, column 0 the PCG signal (decimal float), column 1 the state label (integer 1..4).  Many files are
// parsed by a pool of host threads straight into caller-owned (pinned) staging buffers, from which the caller issues one async
// H2D copy per recording (hss/utils/ingest.py overlaps it with the previous recording's FSST).  Host code only: no CUDA calls.
//
// Numbers go through std::from_chars (correctly rounded decimal -> double) and are narrowed to float exactly as
// torch.tensor(float64 ndarray, dtype=float32) does in the reference.
#include "hssb_common.cuh"
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char *path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); fd = -1; return false; }
        n = (size_t)st.st_size;
        if (n == 0) { p = ""; return true; }
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { ::close(fd); fd = -1; return false; }
        madvise(m, n, MADV_SEQUENTIAL);
        p = static_cast<const char *>(m);
        return true;
    }
    ~Mapped()
    {
        if (p && n) munmap(const_cast<char *>(p), n);
        if (fd >= 0) ::close(fd);
    }
};

inline const char *skip_line(const char *s, const char *end)
{
    const char *nl = static_cast<const char *>(memchr(s, '\n', (size_t)(end - s)));
    return nl ? nl + 1 : end;
}

inline bool blank_line(const char *s, const char *e)
{
    for (; s < e; ++s)
        if (*s != ' ' && *s != '\t' && *s != '\r') return false;
    return true;
}

// data rows of a file: lines after the header that are not blank
int64_t count_rows(const char *s, const char *end)
{
    s = skip_line(s, end);          // header (skiprows=1)
    int64_t rows = 0;
    while (s < end) {
        const char *e = static_cast<const char *>(memchr(s, '\n', (size_t)(end - s)));
        if (!e) e = end;
        if (!blank_line(s, e)) ++rows;
        s = e < end ? e + 1 : end;
    }
    return rows;
}

inline const char *skip_blanks(const char *s, const char *e)
{
    while (s < e && (*s == ' ' || *s == '\t')) ++s;
    return s;
}

// one field: decimal number, optional leading '+', optional surrounding blanks / quotes are not supported (the dataset has none)
inline bool parse_number(const char *&s, const char *e, double &v)
{
    s = skip_blanks(s, e);
    if (s < e && *s == '+') ++s;
    auto r = std::from_chars(s, e, v);
    if (r.ec != std::errc()) return false;
    s = skip_blanks(r.ptr, e);
    return true;
}

// returns rows written, or -(line number) of the first malformed row
int64_t parse_file(const char *s, const char *end, float *x, int64_t *y, int64_t capacity)
{
    s = skip_line(s, end);
    int64_t rows = 0, line = 1;
    while (s < end) {
        ++line;
        const char *e = static_cast<const char *>(memchr(s, '\n', (size_t)(end - s)));
        if (!e) e = end;
        const char *le = e;
        if (le > s && le[-1] == '\r') --le;
        if (!blank_line(s, le)) {
            if (rows >= capacity) return -line;
            double a, b;
            const char *c = s;
            if (!parse_number(c, le, a) || c >= le || *c != ',') return -line;
            ++c;
            if (!parse_number(c, le, b) || (c < le && *c != ',')) return -line;     // further columns are ignored
            x[rows] = (float)a;
            y[rows] = (int64_t)std::llround(b);
            ++rows;
        }
        s = e < end ? e + 1 : end;
    }
    return rows;
}

template <class F>
void run_pool(int n_items, int threads, F fn)
{
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > n_items) threads = n_items;
    std::atomic<int> next{0};
    auto worker = [&] {
        for (int i = next.fetch_add(1); i < n_items; i = next.fetch_add(1)) fn(i);
    };
    if (threads <= 1) { worker(); return; }
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

}  // namespace

// rows_out[i] = number of data rows of paths[i] (header skipped, blank lines ignored)
extern "C" int hssb_csv_scan(const char *const *paths, int n_files, int threads, int64_t *rows_out)
{
    using namespace hssb;
    if (!paths || !rows_out) return fail(HSSB_E_NULL, "hssb_csv_scan: null pointer");
    if (n_files < 0) return fail(HSSB_E_SHAPE, "hssb_csv_scan: n_files=%d", n_files);
    std::atomic<int> bad{-1};
    run_pool(n_files, threads, [&](int i) {
        Mapped f;
        if (!paths[i] || !f.open(paths[i])) { int exp = -1; bad.compare_exchange_strong(exp, i); rows_out[i] = -1; return; }
        rows_out[i] = count_rows(f.p, f.p + f.n);
    });
    if (bad >= 0) return fail(HSSB_E_IO, "hssb_csv_scan: cannot read '%s'", paths[bad] ? paths[bad] : "(null)");
    return 0;
}

// Parses file i into signal[offsets[i] .. offsets[i+1]) / labels[same range]; offsets from the row counts of hssb_csv_scan.
extern "C" int hssb_csv_parse(const char *const *paths, int n_files, int threads, const int64_t *offsets, float *signal, int64_t *labels)
{
    using namespace hssb;
    if (!paths || !offsets || !signal || !labels) return fail(HSSB_E_NULL, "hssb_csv_parse: null pointer");
    if (n_files < 0) return fail(HSSB_E_SHAPE, "hssb_csv_parse: n_files=%d", n_files);
    for (int i = 0; i < n_files; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(HSSB_E_SHAPE, "hssb_csv_parse: offsets must not decrease");
    std::atomic<int> bad{-1};
    std::vector<int64_t> where(n_files > 0 ? n_files : 1, 0);
    run_pool(n_files, threads, [&](int i) {
        Mapped f;
        if (!paths[i] || !f.open(paths[i])) { where[i] = 0; int exp = -1; bad.compare_exchange_strong(exp, i); return; }
        const int64_t cap = offsets[i + 1] - offsets[i];
        const int64_t got = parse_file(f.p, f.p + f.n, signal + offsets[i], labels + offsets[i], cap);
        if (got != cap) { where[i] = got < 0 ? got : -(got + 2); int exp = -1; bad.compare_exchange_strong(exp, i); }
    });
    if (bad >= 0) {
        const int i = bad;
        if (where[i] == 0) return fail(HSSB_E_IO, "hssb_csv_parse: cannot read '%s'", paths[i] ? paths[i] : "(null)");
        return fail(HSSB_E_IO, "hssb_csv_parse: '%s': malformed row or row count differs from the scan (line %lld)", paths[i], (long long)-where[i]);
    }
    return 0;
}
