// Host-side loader for the reference's frame files.
//
// DataStorage.save_raw_data (CodeVector/codevector_classes.py:438-444) writes a recording as a
// JSON list of RawDataMFCC.to_dict() objects (:252-264): 320 raw samples, five scalars, the
// 13-coefficient "mfcc_vector" and the recording name, indent=2 — ~8 KB of text per 20 ms frame,
// of which the hot path needs 13 numbers.  json.load + from_dict (:478-495) builds every object
// (and re-runs librosa on the raw samples); this scanner walks the text once, skips everything
// that is not a key with memchr, and parses only the arrays behind "mfcc_vector" with strtod
// (correctly rounded, so the values are bit-identical to Python's float()).
#include <locale.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

// s points at the opening quote; returns one past the closing quote (nullptr if unterminated)
const char *skip_string(const char *s, const char *end) {
    for (const char *p = s + 1; p < end; ++p) {
        p = static_cast<const char *>(memchr(p, '"', (size_t)(end - p)));
        if (!p) return nullptr;
        // a quote is escaped iff preceded by an odd number of backslashes
        int bs = 0;
        for (const char *q = p - 1; q > s && *q == '\\'; --q) ++bs;
        if ((bs & 1) == 0) return p + 1;
    }
    return nullptr;
}

// the "C" locale for strtod_l: a comma-decimal LC_NUMERIC of the calling process must not change the parse
locale_t c_locale() {
    static locale_t loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    return loc;
}

inline const char *skip_ws(const char *p, const char *end) {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
    return p;
}

}  // namespace

extern "C" int64_t hmmb_frames_json_scan(const char *text, int64_t len, double *mfcc_out, int64_t cap_frames) {
    using namespace hmmb;
    if (!text || len < 0) { set_error("hmmb_frames_json_scan: null text"); return HMMB_ERR_ARG; }
    static const char KEY[] = "\"mfcc_vector\"";
    const size_t KLEN = sizeof(KEY) - 1;
    const char *p = text, *end = text + len;
    int64_t frames = 0;
    while (p < end) {
        p = static_cast<const char *>(memchr(p, '"', (size_t)(end - p)));
        if (!p) break;
        const char *after = skip_string(p, end);
        if (!after) { set_error("frame JSON: unterminated string at byte %lld", (long long)(p - text)); return HMMB_ERR_ARG; }
        const bool is_key_name = (size_t)(after - p) == KLEN && memcmp(p, KEY, KLEN) == 0;
        p = after;
        if (!is_key_name) continue;
        const char *q = skip_ws(p, end);
        if (q >= end || *q != ':') continue;  // a string VALUE that happens to read "mfcc_vector"
        q = skip_ws(q + 1, end);
        if (q >= end || *q != '[') { set_error("frame JSON: mfcc_vector is not an array (frame %lld)", (long long)frames); return HMMB_ERR_ARG; }
        ++q;
        int n = 0;
        for (;;) {
            q = skip_ws(q, end);
            if (q >= end) { set_error("frame JSON: truncated mfcc_vector (frame %lld)", (long long)frames); return HMMB_ERR_ARG; }
            if (*q == ']') { ++q; break; }
            if (*q == ',') { ++q; continue; }
            char *stop = nullptr;
            double v;
            if (end - q >= 3 && memcmp(q, "NaN", 3) == 0) { v = NAN; stop = const_cast<char *>(q) + 3; }            // json.dump spellings
            else if (end - q >= 8 && memcmp(q, "Infinity", 8) == 0) { v = INFINITY; stop = const_cast<char *>(q) + 8; }
            else if (end - q >= 9 && memcmp(q, "-Infinity", 9) == 0) { v = -INFINITY; stop = const_cast<char *>(q) + 9; }
            else {
                // the token is copied (length clamped to the buffer) and NUL-terminated, so a file that ends inside
                // a number is never read past `end`; strtod_l with the "C" locale ignores the process's LC_NUMERIC
                char tok[64];
                size_t len = 0;
                while (len < sizeof(tok) - 1 && q + len < end) {
                    const char ch = q[len];
                    if (!((ch >= '0' && ch <= '9') || ch == '+' || ch == '-' || ch == '.' || ch == 'e' || ch == 'E')) break;
                    tok[len++] = ch;
                }
                tok[len] = 0;
                char *tstop = nullptr;
                v = strtod_l(tok, &tstop, c_locale());
                stop = const_cast<char *>(q) + (tstop - tok);
            }
            if (stop == q) { set_error("frame JSON: bad number in mfcc_vector (frame %lld)", (long long)frames); return HMMB_ERR_ARG; }
            if (n < HMMB_DIM && mfcc_out && frames < cap_frames) mfcc_out[frames * HMMB_DIM + n] = v;
            ++n;
            q = stop;
        }
        if (n != HMMB_DIM) {
            // the reference's distance raises ValueError("Vectors must be of size 13.") for such a frame
            set_error("Vectors must be of size 13. (frame %lld has %d coefficients)", (long long)frames, n);
            return HMMB_ERR_RANGE;
        }
        ++frames;
        p = q;
    }
    return frames;
}
