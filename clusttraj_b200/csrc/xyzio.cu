// Native one-shot trajectory reader: the reference parses every frame again for every row of
// the matrix through OpenBabel (clusttraj/distmat.py:56-61,123-131 + utils.get_mol_info, utils.py:20-31); here the whole
// multi-frame XYZ file becomes (atomic numbers int32[M][N], coordinates float64[M][N][3]) in ONE pass, parsed by all
// host cores: the file is read once, frame starts are indexed by counting newlines (every frame is N + 2 lines), frames
// are parsed in parallel with strtod (correctly rounded, i.e. the bits Python's float() gives).  Anything irregular —
// frames of different sizes, blank lines between frames, short lines, unknown element symbols — is reported as an error
// and the Python side falls back to its general line-by-line parser, which also words the exceptions.
// Host-only code (no kernels): at 20 000 frames x 100 atoms (60 MB of text) 2.7 s of numpy.loadtxt become ~0.1 s, which
// matters once the matrix itself takes 29 ms.
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/clusttraj_b200.h"

namespace ct {
void set_global_error(const std::string& msg);   // engine.cu: what ct_last_error(NULL) returns

static const char* const SYMBOLS[] = {
    "X",  "H",  "He", "Li", "Be", "B",  "C",  "N",  "O",  "F",  "Ne", "Na", "Mg", "Al", "Si", "P",  "S",  "Cl", "Ar", "K",
    "Ca", "Sc", "Ti", "V",  "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr", "Rb", "Sr", "Y",
    "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I",  "Xe", "Cs", "Ba", "La", "Ce", "Pr",
    "Nd", "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W",  "Re", "Os", "Ir", "Pt", "Au",
    "Hg", "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U",  "Np", "Pu", "Am", "Cm", "Bk", "Cf", "Es",
    "Fm", "Md", "No", "Lr", "Rf", "Db", "Sg", "Bh", "Hs", "Mt", "Ds", "Rg", "Cn", "Nh", "Fl", "Mc", "Lv", "Ts", "Og"};
constexpr int N_SYMBOLS = sizeof(SYMBOLS) / sizeof(SYMBOLS[0]);

static int lookup_symbol(const char* s, int len) {   // case-insensitive exact match, -1 if unknown
    if (len < 1 || len > 2) return -1;
    for (int z = 0; z < N_SYMBOLS; ++z) {
        const char* t = SYMBOLS[z];
        const int tl = t[1] ? 2 : 1;
        if (tl != len) continue;
        if (std::tolower((unsigned char)s[0]) != std::tolower((unsigned char)t[0])) continue;
        if (len == 2 && std::tolower((unsigned char)s[1]) != std::tolower((unsigned char)t[1])) continue;
        return z;
    }
    return -1;
}

// clusttraj_b200/xyz.py:_atomic_number: an integer token is the atomic number; else the whole token, then its first two,
// then its first character, as an element symbol
static int atomic_number(const char* tok, int len) {
    bool digits = len > 0;
    for (int k = 0; k < len; ++k) digits = digits && tok[k] >= '0' && tok[k] <= '9';
    if (digits) {
        long v = 0;
        for (int k = 0; k < len && v < 100000; ++k) v = v * 10 + (tok[k] - '0');
        return (int)v;
    }
    int z = lookup_symbol(tok, len);
    if (z < 0) {   // Python: Z_OF.get(token[:2]) or Z_OF.get(token[:1])  (a falsy first hit falls through to the second)
        const int z2 = lookup_symbol(tok, len < 2 ? len : 2);
        z = z2 > 0 ? z2 : lookup_symbol(tok, 1);
    }
    return z;
}

struct XyzFile {
    std::vector<char> text;          // whole file + '\0'
    std::vector<size_t> frame_start; // offset of every frame's first line, + end sentinel
    long n_atoms = 0;
};

static inline const char* line_end(const char* p, const char* end) {
    const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
    return q ? q : end;
}

static int load_and_index(const char* path, XyzFile& f, std::string& err) {
    FILE* fh = fopen(path, "rb");
    if (!fh) { err = std::string("cannot open ") + path; return CT_ERR_ARG; }
    fseek(fh, 0, SEEK_END);
    const long size = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    if (size <= 0) { fclose(fh); err = "empty file"; return CT_ERR_ARG; }
    f.text.resize((size_t)size + 1);
    const size_t got = fread(f.text.data(), 1, (size_t)size, fh);
    fclose(fh);
    if (got != (size_t)size) { err = "short read"; return CT_ERR_ARG; }
    f.text[(size_t)size] = '\0';
    const char* base = f.text.data();
    const char* end = base + size;
    while (end > base && (end[-1] == '\n' || end[-1] == '\r' || end[-1] == ' ' || end[-1] == '\t')) --end;   // trailing blank lines
    char* stop = nullptr;
    f.n_atoms = strtol(base, &stop, 10);
    if (stop == base || f.n_atoms < 1 || f.n_atoms > 100000000L) { err = "first line is not an atom count"; return CT_ERR_ARG; }
    const long per_frame = f.n_atoms + 2;
    long line = 0;
    const char* p = base;
    while (p < end) {
        if (line % per_frame == 0) f.frame_start.push_back((size_t)(p - base));
        const char* q = line_end(p, end);
        ++line;
        p = q < end ? q + 1 : end;
    }
    if (line % per_frame != 0) { err = "the file is not a whole number of (atom count + 2)-line frames"; return CT_ERR_ARG; }
    f.frame_start.push_back((size_t)(end - base));
    return CT_OK;
}

static bool parse_frame(const XyzFile& f, size_t fr, double* xyz, int32_t* zs, std::string& err) {
    const char* base = f.text.data();
    const char* p = base + f.frame_start[fr];
    const char* end = base + f.frame_start[fr + 1];
    char* stop = nullptr;
    const long n = strtol(p, &stop, 10);
    if (stop == p || n != f.n_atoms) { err = "frame " + std::to_string(fr) + ": atom count differs from the first frame's"; return false; }
    p = line_end(p, end); if (p < end) ++p;       // count line
    p = line_end(p, end); if (p < end) ++p;       // comment line
    for (long k = 0; k < f.n_atoms; ++k) {
        if (p >= end) { err = "frame " + std::to_string(fr) + ": truncated"; return false; }
        const char* eol = line_end(p, end);
        const char* t = p;
        while (t < eol && (*t == ' ' || *t == '\t' || *t == '\r')) ++t;
        const char* t1 = t;
        while (t1 < eol && !(*t1 == ' ' || *t1 == '\t' || *t1 == '\r')) ++t1;
        const int z = atomic_number(t, (int)(t1 - t));
        if (z < 0) { err = "frame " + std::to_string(fr) + ": unknown element symbol"; return false; }
        zs[k] = z;
        const char* c = t1;
        for (int a = 0; a < 3; ++a) {
            while (c < eol && (*c == ' ' || *c == '\t' || *c == '\r')) ++c;
            if (c >= eol) { err = "frame " + std::to_string(fr) + ": an atom line has fewer than four fields"; return false; }
            const double v = strtod(c, &stop);
            // the token must end at white space: anything float() would refuse goes to the general parser
            if (stop == c || stop > eol || (stop < eol && !(*stop == ' ' || *stop == '\t' || *stop == '\r'))) {
                err = "frame " + std::to_string(fr) + ": a coordinate is not a plain number";
                return false;
            }
            xyz[3 * k + a] = v;
            c = stop;
        }
        p = eol < end ? eol + 1 : end;
    }
    return true;
}

// ---- PDB (MODEL ... ENDMDL frames, fixed columns) and GRO (title, count, fixed-column atom lines in nm, box) -----------
// Same two rules as clusttraj_b200/xyz.py:read_pdb / read_gro, which remain the fall-back for anything refused here.

static int load_text(const char* path, std::vector<char>& text, std::string& err) {
    FILE* fh = fopen(path, "rb");
    if (!fh) { err = std::string("cannot open ") + path; return CT_ERR_ARG; }
    fseek(fh, 0, SEEK_END);
    const long size = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    if (size <= 0) { fclose(fh); err = "empty file"; return CT_ERR_ARG; }
    text.resize((size_t)size + 1);
    const size_t got = fread(text.data(), 1, (size_t)size, fh);
    fclose(fh);
    if (got != (size_t)size) { err = "short read"; return CT_ERR_ARG; }
    text[(size_t)size] = '\0';
    return CT_OK;
}

// float(line[a:b]) for a fixed-column field: the whole field, blanks stripped, must be one number
static bool field_number(const char* line, const char* eol, int a, int b, double& v) {
    const char* lo = line + a < eol ? line + a : eol;
    const char* hi = line + b < eol ? line + b : eol;
    while (lo < hi && (*lo == ' ' || *lo == '\t' || *lo == '\r')) ++lo;
    while (hi > lo && (hi[-1] == ' ' || hi[-1] == '\t' || hi[-1] == '\r')) --hi;
    if (lo >= hi || hi - lo > 63) return false;
    char buf[64];
    memcpy(buf, lo, (size_t)(hi - lo));
    buf[hi - lo] = '\0';
    char* stop = nullptr;
    v = strtod(buf, &stop);
    return stop == buf + (hi - lo);
}

static int letters_of(const char* line, const char* eol, int a, int b, char* out) {   // the alphabetic characters of line[a:b]
    int n = 0;
    for (const char* c = line + a; c < line + b && c < eol; ++c)
        if (std::isalpha((unsigned char)*c)) out[n++] = *c;
    out[n] = '\0';
    return n;
}

struct Frames {
    std::vector<char> text;
    std::vector<size_t> start, stop;   // byte range of every frame
    long n_atoms = 0;
};

// PDB: a frame ends at a record starting with "END" (ENDMDL, END) or at the end of the file; only frames with atoms count
static int index_pdb(Frames& f, std::string& err) {
    const char* base = f.text.data();
    const char* end = base + f.text.size() - 1;
    const char* p = base;
    size_t frame_begin = 0;
    long atoms = 0;
    auto flush = [&](size_t upto) {
        if (atoms > 0) {
            if (f.n_atoms == 0) f.n_atoms = atoms;
            if (atoms != f.n_atoms) return false;
            f.start.push_back(frame_begin);
            f.stop.push_back(upto);
        }
        atoms = 0;
        frame_begin = upto;
        return true;
    };
    while (p < end) {
        const char* q = line_end(p, end);
        if (q - p >= 6 && (!memcmp(p, "ATOM  ", 6) || !memcmp(p, "HETATM", 6))) ++atoms;
        else if (q - p >= 3 && !memcmp(p, "END", 3)) {
            if (!flush((size_t)((q < end ? q + 1 : end) - base))) { err = "frames have different atom counts"; return CT_ERR_ARG; }
        }
        p = q < end ? q + 1 : end;
    }
    if (!flush((size_t)(end - base))) { err = "frames have different atom counts"; return CT_ERR_ARG; }
    if (f.start.empty()) { err = "no ATOM/HETATM records"; return CT_ERR_ARG; }
    return CT_OK;
}

static bool parse_pdb_frame(const Frames& f, size_t fr, double* xyz, int32_t* zs, std::string& err) {
    const char* base = f.text.data();
    const char* p = base + f.start[fr];
    const char* end = base + f.stop[fr];
    long k = 0;
    while (p < end) {
        const char* q = line_end(p, end);
        if (q - p >= 6 && (!memcmp(p, "ATOM  ", 6) || !memcmp(p, "HETATM", 6))) {
            const char* eol = q;
            while (eol > p && eol[-1] == '\r') --eol;
            char el[8];
            int n = 0;
            for (const char* c = p + 76; c < p + 78 && c < eol; ++c)           // line[76:78].strip()
                if (!(*c == ' ' || *c == '\t')) el[n++] = *c;
            el[n] = '\0';
            int z;
            if (n > 0) z = atomic_number(el, n);
            else {
                // no element columns: the alignment convention of the atom-name field, clusttraj_b200/xyz.py:pdb_element_from_name
                char name[5] = {' ', ' ', ' ', ' ', '\0'};
                for (int t = 0; t < 4 && p + 12 + t < eol; ++t) name[t] = p[12 + t];
                char sym[3] = {0, 0, 0};
                int m = 0;
                if (std::isalpha((unsigned char)name[0])) {
                    const bool four = name[1] != ' ' && name[2] != ' ' && name[3] != ' ';
                    sym[0] = name[0]; sym[1] = name[1];
                    if (std::isalpha((unsigned char)name[1]) && lookup_symbol(sym, 2) >= 0 && !((name[0] == 'H' || name[0] == 'h') && four)) m = 2;
                    else m = 1;
                } else {
                    for (int t = 1; t < 4 && !m; ++t)
                        if (std::isalpha((unsigned char)name[t])) { sym[0] = name[t]; m = 1; }
                }
                z = m ? atomic_number(sym, m) : -1;
            }
            if (z < 0) { err = "frame " + std::to_string(fr) + ": unknown element"; return false; }
            if (!field_number(p, eol, 30, 38, xyz[3 * k]) || !field_number(p, eol, 38, 46, xyz[3 * k + 1]) ||
                !field_number(p, eol, 46, 54, xyz[3 * k + 2])) {
                err = "frame " + std::to_string(fr) + ": bad coordinate field";
                return false;
            }
            zs[k++] = z;
        }
        p = q < end ? q + 1 : end;
    }
    return k == f.n_atoms;
}

// GRO: title line, atom count, N fixed-column atom lines, box line; two consecutive blank lines are skipped one at a time
static int index_gro(Frames& f, std::string& err) {
    const char* base = f.text.data();
    const char* end = base + f.text.size() - 1;
    std::vector<size_t> line_at;
    for (const char* p = base; p < end;) { line_at.push_back((size_t)(p - base)); const char* q = line_end(p, end); p = q < end ? q + 1 : end; }
    if (end > base && end[-1] == '\n') line_at.push_back((size_t)(end - base));   // Python's split("\n") yields a last empty line
    const size_t nl = line_at.size();
    auto blank = [&](size_t i) {
        const char* a = base + line_at[i];
        const char* b = i + 1 < nl ? base + line_at[i + 1] : end;
        for (const char* c = a; c < b; ++c) if (!(*c == ' ' || *c == '\t' || *c == '\r' || *c == '\n')) return false;
        return true;
    };
    size_t pos = 0;
    while (pos + 1 < nl) {
        if (blank(pos) && blank(pos + 1)) { ++pos; continue; }
        const char* c = base + line_at[pos + 1];
        char* stop = nullptr;
        const long n = strtol(c, &stop, 10);
        if (stop == c || n < 1) { err = "bad atom count line"; return CT_ERR_ARG; }
        if (pos + 2 + (size_t)n > nl) {
            err = "truncated frame";
            return CT_ERR_ARG;
        }
        if (f.n_atoms == 0) f.n_atoms = n;
        if (n != f.n_atoms) { err = "frames have different atom counts"; return CT_ERR_ARG; }
        f.start.push_back(line_at[pos + 2]);
        f.stop.push_back(pos + 2 + (size_t)n < nl ? line_at[pos + 2 + (size_t)n] : (size_t)(end - base));
        pos += 3 + (size_t)n;
    }
    if (f.start.empty()) { err = "no frames"; return CT_ERR_ARG; }
    return CT_OK;
}

static bool parse_gro_frame(const Frames& f, size_t fr, double* xyz, int32_t* zs, std::string& err) {
    const char* base = f.text.data();
    const char* p = base + f.start[fr];
    const char* end = base + f.stop[fr];
    for (long k = 0; k < f.n_atoms; ++k) {
        if (p >= end) { err = "frame " + std::to_string(fr) + ": truncated"; return false; }
        const char* q = line_end(p, end);
        const char* eol = q;
        while (eol > p && eol[-1] == '\r') --eol;
        char name[8];
        const int m = letters_of(p, eol, 10, 15, name);
        // name[:2] if it is an element, has a second letter and that letter is lower case, else name[:1]
        int len = (m > 1 && std::islower((unsigned char)name[1]) && lookup_symbol(name, 2) >= 0) ? 2 : (m > 0 ? 1 : 0);
        const int z = atomic_number(name, len);
        if (z < 0 || len == 0) { err = "frame " + std::to_string(fr) + ": unknown element"; return false; }
        double v[3];
        if (!field_number(p, eol, 20, 28, v[0]) || !field_number(p, eol, 28, 36, v[1]) || !field_number(p, eol, 36, 44, v[2])) {
            err = "frame " + std::to_string(fr) + ": bad coordinate field";
            return false;
        }
        zs[k] = z;
        xyz[3 * k] = 10.0 * v[0]; xyz[3 * k + 1] = 10.0 * v[1]; xyz[3 * k + 2] = 10.0 * v[2];   // nm -> Angstrom
        p = q < end ? q + 1 : end;
    }
    return true;
}

template <class Index, class Parse>
static int read_frames(const char* what, const char* path, int64_t* n_frames, int32_t* n_atoms, double* coords,
                       int32_t* atomic_numbers, int64_t capacity_frames, int32_t n_threads, Index index, Parse parse) {
    if (!path || !n_frames || !n_atoms) { set_global_error(std::string(what) + ": NULL argument"); return CT_ERR_ARG; }
    Frames f;
    std::string err;
    int rc = load_text(path, f.text, err);
    if (rc == CT_OK) rc = index(f, err);
    if (rc != CT_OK) { set_global_error(std::string(what) + ": " + err); return rc; }
    const int64_t M = (int64_t)f.start.size();
    *n_frames = M;
    *n_atoms = (int32_t)f.n_atoms;
    if (!coords || !atomic_numbers) return CT_OK;   // size query
    if (capacity_frames < M) { set_global_error(std::string(what) + ": output arrays are too small"); return CT_ERR_ARG; }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 64 ? 64 : nt);
    if ((int64_t)nt > M) nt = (int)M;
    std::vector<std::string> errs((size_t)nt);
    std::vector<char> failed((size_t)nt, 0);
    const size_t N = (size_t)f.n_atoms;
    auto work = [&](int t) {
        const int64_t lo = M * t / nt, hi = M * (t + 1) / nt;
        for (int64_t fr = lo; fr < hi; ++fr)
            if (!parse(f, (size_t)fr, coords + (size_t)fr * N * 3, atomic_numbers + (size_t)fr * N, errs[(size_t)t])) { failed[(size_t)t] = 1; return; }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < nt; ++t)
        if (failed[(size_t)t]) { set_global_error(std::string(what) + ": " + (errs[(size_t)t].empty() ? "malformed frame" : errs[(size_t)t])); return CT_ERR_ARG; }
    return CT_OK;
}

}  // namespace ct

extern "C" {

int ct_read_pdb(const char* path, int64_t* n_frames, int32_t* n_atoms, double* coords, int32_t* atomic_numbers,
                int64_t capacity_frames, int32_t n_threads) {
    return ct::read_frames("ct_read_pdb", path, n_frames, n_atoms, coords, atomic_numbers, capacity_frames, n_threads,
                           ct::index_pdb, ct::parse_pdb_frame);
}

int ct_read_gro(const char* path, int64_t* n_frames, int32_t* n_atoms, double* coords, int32_t* atomic_numbers,
                int64_t capacity_frames, int32_t n_threads) {
    return ct::read_frames("ct_read_gro", path, n_frames, n_atoms, coords, atomic_numbers, capacity_frames, n_threads,
                           ct::index_gro, ct::parse_gro_frame);
}

int ct_read_xyz(const char* path, int64_t* n_frames, int32_t* n_atoms, double* coords, int32_t* atomic_numbers,
                int64_t capacity_frames, int32_t n_threads) {
    if (!path || !n_frames || !n_atoms) { ct::set_global_error("ct_read_xyz: NULL argument"); return CT_ERR_ARG; }
    ct::XyzFile f;
    std::string err;
    int rc = ct::load_and_index(path, f, err);
    if (rc != CT_OK) { ct::set_global_error("ct_read_xyz: " + err); return rc; }
    const int64_t M = (int64_t)f.frame_start.size() - 1;
    *n_frames = M;
    *n_atoms = (int32_t)f.n_atoms;
    if (!coords || !atomic_numbers) return CT_OK;   // size query
    if (capacity_frames < M) { ct::set_global_error("ct_read_xyz: output arrays are too small"); return CT_ERR_ARG; }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if ((int64_t)nt > M) nt = (int)M;
    std::vector<std::string> errs((size_t)nt);
    std::vector<std::thread> pool;
    const size_t N = (size_t)f.n_atoms;
    auto work = [&](int t) {
        const int64_t lo = M * t / nt, hi = M * (t + 1) / nt;
        for (int64_t fr = lo; fr < hi; ++fr)
            if (!ct::parse_frame(f, (size_t)fr, coords + (size_t)fr * N * 3, atomic_numbers + (size_t)fr * N, errs[(size_t)t])) return;
    };
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (const auto& e : errs)
        if (!e.empty()) { ct::set_global_error("ct_read_xyz: " + e); return CT_ERR_ARG; }
    return CT_OK;
}

// Text of F consecutive XYZ frames in the layout the reference emits through pybel.Outputfile("xyz", ...)
// (clusttraj/io.py:870-899,1000-1015): "<N>\n<title>\n" then "%-3s%15.5f%15.5f%15.5f\n" per atom.  Frames are formatted
// by n_threads host threads into one malloc'ed buffer (*out, *length bytes; release it with ct_buffer_free).
int ct_format_xyz(const int32_t* atomic_numbers, const double* coords, int64_t n_frames, int32_t n_atoms, const char* titles,
                  char** out, int64_t* length, int32_t n_threads) {
    if (!atomic_numbers || !coords || !out || !length || n_frames < 0 || n_atoms < 1) {
        ct::set_global_error("ct_format_xyz: bad argument");
        return CT_ERR_ARG;
    }
    for (int k = 0; k < n_atoms; ++k)
        if (atomic_numbers[k] < 0 || atomic_numbers[k] >= ct::N_SYMBOLS) {
            ct::set_global_error("ct_format_xyz: atomic number out of range");
            return CT_ERR_ARG;
        }
    std::vector<const char*> title((size_t)n_frames, "");
    if (titles) {
        const char* t = titles;
        for (int64_t f = 0; f < n_frames; ++f) { title[(size_t)f] = t; t += strlen(t) + 1; }
    }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if ((int64_t)nt > n_frames) nt = n_frames > 0 ? (int)n_frames : 1;
    std::vector<std::string> part((size_t)nt);
    auto work = [&](int t) {
        const int64_t lo = n_frames * t / nt, hi = n_frames * (t + 1) / nt;
        std::string& s = part[(size_t)t];
        s.reserve((size_t)(hi - lo) * ((size_t)n_atoms * 49 + 64));
        char line[1100];
        for (int64_t f = lo; f < hi; ++f) {
            s += std::to_string(n_atoms);
            s += '\n';
            s += title[(size_t)f];
            s += '\n';
            const double* x = coords + (size_t)f * n_atoms * 3;
            for (int k = 0; k < n_atoms; ++k) {
                const int len = snprintf(line, sizeof line, "%-3s%15.5f%15.5f%15.5f\n", ct::SYMBOLS[atomic_numbers[k]], x[3 * k],
                                         x[3 * k + 1], x[3 * k + 2]);
                s.append(line, (size_t)(len < (int)sizeof line ? len : (int)sizeof line - 1));
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    size_t total = 0;
    for (const auto& s : part) total += s.size();
    char* buf = (char*)malloc(total ? total : 1);
    if (!buf) { ct::set_global_error("ct_format_xyz: out of memory"); return CT_ERR_ARG; }
    size_t pos = 0;
    for (const auto& s : part) { memcpy(buf + pos, s.data(), s.size()); pos += s.size(); }
    *out = buf;
    *length = (int64_t)total;
    return CT_OK;
}

void ct_buffer_free(void* p) { free(p); }

}  // extern "C"
