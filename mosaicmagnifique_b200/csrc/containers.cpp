// See containers.h. Host-only code: stream parsing, a PNG codec (inflate per RFC 1951, zlib framing per RFC 1950, PNG chunks,
// filters and CRC per the PNG specification) and the two container layouts of the reference.
#include "containers.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>

namespace mm {

namespace {

constexpr uint32_t kMaxSide = 65536;            // images in these containers are cell masks and library tiles
constexpr size_t kMaxInflated = (size_t)1 << 31;  // refuse to inflate more than 2 GiB from one stream

// ------------------------------------------------------------------ checksums

uint32_t crc32_of(const uint8_t *p, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k)
                c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i)
        crc = table[(crc ^ p[i]) & 255] ^ (crc >> 8);
    return ~crc;
}

uint32_t adler32_of(const uint8_t *p, size_t n)
{
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; ++i) {
        a = (a + p[i]) % 65521u;
        b = (b + a) % 65521u;
    }
    return (b << 16) | a;
}

// ------------------------------------------------------------------ inflate (RFC 1951)

struct BitReader {
    const uint8_t *p;
    size_t n, pos = 0;
    uint32_t buf = 0;
    int cnt = 0;
    bool overrun = false;
    uint32_t bits(int need)
    {
        while (cnt < need) {
            if (pos >= n) {
                overrun = true;
                return 0;
            }
            buf |= (uint32_t)p[pos++] << cnt;
            cnt += 8;
        }
        const uint32_t v = need ? buf & ((1u << need) - 1) : 0;
        buf >>= need;
        cnt -= need;
        return v;
    }
    void align_to_byte()
    {
        buf = 0;
        cnt = 0;
    }
};

// canonical Huffman code: count[len] codes of each length, symbols sorted by (length, value)
struct Huffman {
    uint16_t count[16];
    uint16_t symbol[288];
    bool build(const uint8_t *lengths, int n)
    {
        memset(count, 0, sizeof count);
        for (int i = 0; i < n; ++i)
            count[lengths[i]]++;
        int left = 1;
        for (int len = 1; len < 16; ++len) {
            left = (left << 1) - count[len];
            if (left < 0)
                return false;  // over-subscribed
        }
        uint16_t offs[16];
        offs[1] = 0;
        for (int len = 1; len < 15; ++len)
            offs[len + 1] = (uint16_t)(offs[len] + count[len]);
        for (int i = 0; i < n; ++i)
            if (lengths[i])
                symbol[offs[lengths[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(BitReader &br) const
    {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; ++len) {
            code |= (int)br.bits(1);
            if (br.overrun)
                return -1;
            const int c = count[len];
            if (code - c < first)
                return symbol[index + (code - first)];
            index += c;
            first = (first + c) << 1;
            code <<= 1;
        }
        return -1;
    }
};

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint16_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                4097, 6145, 8193, 12289, 16385, 24577};
const uint16_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

bool inflate_codes(BitReader &br, const Huffman &lit, const Huffman &dist, std::vector<uint8_t> &out)
{
    for (;;) {
        const int sym = lit.decode(br);
        if (sym < 0)
            return false;
        if (out.size() > kMaxInflated)
            return false;
        if (sym < 256) {
            out.push_back((uint8_t)sym);
        } else if (sym == 256) {
            return true;
        } else {
            const int li = sym - 257;
            if (li >= 29)
                return false;
            const size_t len = kLenBase[li] + br.bits(kLenExtra[li]);
            const int ds = dist.decode(br);
            if (ds < 0 || ds >= 30)
                return false;
            const size_t d = kDistBase[ds] + br.bits(kDistExtra[ds]);
            if (br.overrun || d > out.size())
                return false;
            const size_t from = out.size() - d;
            for (size_t i = 0; i < len; ++i)
                out.push_back(out[from + i]);
        }
    }
}

bool inflate_raw(const uint8_t *data, size_t n, std::vector<uint8_t> &out)
{
    BitReader br{data, n};
    for (;;) {
        const uint32_t last = br.bits(1), type = br.bits(2);
        if (br.overrun)
            return false;
        if (type == 0) {
            br.align_to_byte();
            if (br.pos + 4 > n)
                return false;
            const uint32_t len = data[br.pos] | (data[br.pos + 1] << 8), nlen = data[br.pos + 2] | (data[br.pos + 3] << 8);
            br.pos += 4;
            if ((len ^ 0xFFFFu) != nlen || br.pos + len > n || out.size() > kMaxInflated)
                return false;
            out.insert(out.end(), data + br.pos, data + br.pos + len);
            br.pos += len;
        } else if (type == 1) {
            uint8_t l[288];
            for (int i = 0; i < 144; ++i) l[i] = 8;
            for (int i = 144; i < 256; ++i) l[i] = 9;
            for (int i = 256; i < 280; ++i) l[i] = 7;
            for (int i = 280; i < 288; ++i) l[i] = 8;
            uint8_t dl[30];
            for (int i = 0; i < 30; ++i) dl[i] = 5;
            Huffman lit, dist;
            lit.build(l, 288);
            dist.build(dl, 30);
            if (!inflate_codes(br, lit, dist, out))
                return false;
        } else if (type == 2) {
            const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
            if (br.overrun || nlen > 286 || ndist > 30)
                return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t lengths[320];
            memset(lengths, 0, sizeof lengths);
            for (int i = 0; i < ncode; ++i)
                lengths[order[i]] = (uint8_t)br.bits(3);
            Huffman lencode;
            if (!lencode.build(lengths, 19))
                return false;
            uint8_t all[320];
            int idx = 0;
            while (idx < nlen + ndist) {
                const int sym = lencode.decode(br);
                if (sym < 0)
                    return false;
                if (sym < 16) {
                    all[idx++] = (uint8_t)sym;
                } else {
                    uint8_t prev = 0;
                    int rep;
                    if (sym == 16) {
                        if (idx == 0)
                            return false;
                        prev = all[idx - 1];
                        rep = 3 + (int)br.bits(2);
                    } else if (sym == 17) {
                        rep = 3 + (int)br.bits(3);
                    } else {
                        rep = 11 + (int)br.bits(7);
                    }
                    if (br.overrun || idx + rep > nlen + ndist)
                        return false;
                    while (rep--)
                        all[idx++] = prev;
                }
            }
            if (all[256] == 0)
                return false;
            Huffman lit, dist;
            if (!lit.build(all, nlen) || !dist.build(all + nlen, ndist))
                return false;
            if (!inflate_codes(br, lit, dist, out))
                return false;
        } else {
            return false;
        }
        if (last)
            return true;
    }
}

bool zlib_decompress(const uint8_t *data, size_t n, std::vector<uint8_t> &out)
{
    if (n < 6 || (data[0] & 15) != 8 || ((data[0] << 8) | data[1]) % 31 != 0 || (data[1] & 32))
        return false;
    if (!inflate_raw(data + 2, n - 6, out))
        return false;
    const uint32_t want = ((uint32_t)data[n - 4] << 24) | (data[n - 3] << 16) | (data[n - 2] << 8) | data[n - 1];
    return adler32_of(out.data(), out.size()) == want;
}

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }
void put_be32(std::vector<uint8_t> &v, uint32_t x)
{
    for (int s = 24; s >= 0; s -= 8)
        v.push_back((uint8_t)(x >> s));
}

int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

// ------------------------------------------------------------------ PNG

bool png_decode(const uint8_t *data, size_t n, Image8 &out, std::string &err)
{
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (n < 8 || memcmp(data, sig, 8) != 0) {
        err = "not a PNG stream";
        return false;
    }
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool end = false;
    while (!end && pos + 12 <= n) {
        const uint32_t len = be32(data + pos);
        const uint8_t *type = data + pos + 4, *body = data + pos + 8;
        if (pos + 12 + (size_t)len > n) {
            err = "truncated PNG chunk";
            return false;
        }
        if (crc32_of(type, 4 + (size_t)len) != be32(body + len)) {
            err = "PNG chunk CRC mismatch";
            return false;
        }
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(body);
            h = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!w || !h || ctype < 0 || idat.empty()) {
        err = "PNG without IHDR / IDAT";
        return false;
    }
    if (interlace) {
        err = "interlaced PNG is not supported";
        return false;
    }
    if (w > kMaxSide || h > kMaxSide) {  // keeps every size product below far from overflow
        err = "PNG larger than 65536 pixels on a side";
        return false;
    }
    const int samples = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depth_ok = depth == 8 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4));
    if (!samples || !depth_ok) {
        err = "PNG colour type / bit depth not supported (8-bit grey, RGB, RGBA, palette)";
        return false;
    }
    std::vector<uint8_t> raw;
    if (!zlib_decompress(idat.data(), idat.size(), raw)) {
        err = "PNG data does not inflate";
        return false;
    }
    const size_t row_bytes = ((size_t)w * samples * depth + 7) / 8, bpp = std::max<size_t>(1, (size_t)samples * depth / 8);
    if (raw.size() < (row_bytes + 1) * h) {
        err = "PNG data too short";
        return false;
    }
    // undo the row filters in place
    std::vector<uint8_t> img(row_bytes * h);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t f = raw[y * (row_bytes + 1)];
        const uint8_t *src = &raw[y * (row_bytes + 1) + 1];
        uint8_t *cur = &img[y * row_bytes];
        const uint8_t *up = y ? &img[(y - 1) * row_bytes] : nullptr;
        for (size_t x = 0; x < row_bytes; ++x) {
            const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v = src[x];
            switch (f) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: err = "bad PNG filter"; return false;
            }
            cur[x] = (uint8_t)v;
        }
    }
    // to OpenCV channel order (what cv::imdecode(IMREAD_UNCHANGED) returns): grey, BGR, BGRA
    out.rows = (int)h;
    out.cols = (int)w;
    out.channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 3 : 4;
    out.px.assign((size_t)w * h * out.channels, 0);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t *r = &img[y * row_bytes];
        uint8_t *d = &out.px[(size_t)y * w * out.channels];
        for (uint32_t x = 0; x < w; ++x) {
            auto sample = [&](uint32_t i) -> int {  // i-th sample of the row at the stored bit depth
                if (depth == 8)
                    return r[i];
                const int per = 8 / depth, shift = (per - 1 - (int)(i % per)) * depth;
                return (r[i / per] >> shift) & ((1 << depth) - 1);
            };
            if (ctype == 0) {
                const int v = sample(x);
                d[x] = (uint8_t)(depth == 8 ? v : v * 255 / ((1 << depth) - 1));
            } else if (ctype == 2) {
                d[3 * x] = r[3 * x + 2]; d[3 * x + 1] = r[3 * x + 1]; d[3 * x + 2] = r[3 * x];
            } else if (ctype == 3) {
                const size_t i = (size_t)sample(x) * 3;
                if (i + 2 < plte.size()) { d[3 * x] = plte[i + 2]; d[3 * x + 1] = plte[i + 1]; d[3 * x + 2] = plte[i]; }
            } else if (ctype == 4) {
                d[4 * x] = d[4 * x + 1] = d[4 * x + 2] = r[2 * x]; d[4 * x + 3] = r[2 * x + 1];
            } else {
                d[4 * x] = r[4 * x + 2]; d[4 * x + 1] = r[4 * x + 1]; d[4 * x + 2] = r[4 * x]; d[4 * x + 3] = r[4 * x + 3];
            }
        }
    }
    return true;
}

void png_encode(const Image8 &img, std::vector<uint8_t> &out)
{
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    out.assign(sig, sig + 8);
    auto chunk = [&](const char *type, const std::vector<uint8_t> &body) {
        put_be32(out, (uint32_t)body.size());
        const size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), body.begin(), body.end());
        put_be32(out, crc32_of(&out[start], out.size() - start));
    };
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)img.cols);
    put_be32(ihdr, (uint32_t)img.rows);
    ihdr.push_back(8);
    ihdr.push_back(img.channels == 1 ? 0 : img.channels == 3 ? 2 : 6);
    ihdr.push_back(0);
    ihdr.push_back(0);
    ihdr.push_back(0);
    chunk("IHDR", ihdr);
    // scanlines in PNG sample order (RGB / RGBA), each with the filter (None / Sub / Up / Average / Paeth) that minimises the sum
    // of absolute signed residuals -- libpng's default heuristic
    const size_t row = (size_t)img.cols * img.channels;
    const int bpp = img.channels;
    std::vector<uint8_t> cur(row), prev(row, 0), cand(row), best(row);
    std::vector<uint8_t> raw;
    raw.reserve((row + 1) * img.rows);
    for (int y = 0; y < img.rows; ++y) {
        const uint8_t *s = &img.px[(size_t)y * row];
        for (int x = 0; x < img.cols; ++x)
            for (int c = 0; c < img.channels; ++c)
                cur[(size_t)x * img.channels + c] = s[(size_t)x * img.channels + (img.channels >= 3 && c < 3 ? 2 - c : c)];
        int best_f = 0;
        uint64_t best_cost = ~0ull;
        for (int f = 0; f < 5; ++f) {
            uint64_t cost = 0;
            for (size_t i = 0; i < row; ++i) {
                const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
                int pred = 0;
                if (f == 1)
                    pred = a;
                else if (f == 2)
                    pred = b;
                else if (f == 3)
                    pred = (a + b) >> 1;
                else if (f == 4) {
                    const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                }
                const uint8_t r = (uint8_t)(cur[i] - pred);
                cand[i] = r;
                cost += r < 128 ? r : 256 - r;
            }
            if (cost < best_cost) {
                best_cost = cost;
                best_f = f;
                best.swap(cand);
            }
        }
        raw.push_back((uint8_t)best_f);
        raw.insert(raw.end(), best.begin(), best.end());
        prev = cur;
    }
    // zlib stream: one deflate block (RFC 1951) -- LZ77 over a 32 KB window with hash chains, then whichever of the fixed and a
    // dynamic Huffman coding of the token stream is shorter (masks: long matches, either is tiny; photographs: almost only literal
    // filter residuals, which only a dynamic code shortens)
    std::vector<uint8_t> z = {0x78, 0x5E};
    {
        static const int len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const int len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const int dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const int dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        struct Token {
            uint16_t len;   // 0 = literal
            uint16_t dist;  // literal value when len == 0
        };
        // ---- LZ77
        std::vector<Token> tokens;
        const size_t n = raw.size();
        tokens.reserve(n / 2 + 16);
        constexpr int kHashBits = 15, kWindow = 32768, kMaxChain = 48;
        std::vector<int32_t> head((size_t)1 << kHashBits, -1), chain(n, -1);
        auto hash3 = [&](size_t i) { return ((raw[i] << 10) ^ (raw[i + 1] << 5) ^ raw[i + 2]) & ((1 << kHashBits) - 1); };
        for (size_t i = 0; i < n;) {
            int best_len = 0, best_dist = 0;
            if (i + 3 <= n) {
                int32_t cand_pos = head[hash3(i)];
                for (int tries = 0; cand_pos >= 0 && (i - (size_t)cand_pos) <= (size_t)kWindow && tries < kMaxChain; ++tries) {
                    const size_t max_len = std::min<size_t>(258, n - i);
                    size_t l = 0;
                    while (l < max_len && raw[cand_pos + l] == raw[i + l])
                        ++l;
                    if ((int)l > best_len) {
                        best_len = (int)l;
                        best_dist = (int)(i - (size_t)cand_pos);
                        if (l == max_len)
                            break;
                    }
                    cand_pos = chain[cand_pos];
                }
            }
            const size_t step = best_len >= 3 ? (size_t)best_len : 1;
            tokens.push_back(best_len >= 3 ? Token{(uint16_t)best_len, (uint16_t)best_dist} : Token{0, raw[i]});
            for (size_t k = i; k < i + step; ++k)  // enter the covered positions into the hash chains
                if (k + 3 <= n) {
                    const int h = hash3(k);
                    chain[k] = head[h];
                    head[h] = (int32_t)k;
                }
            i += step;
        }
        auto len_code = [&](int len) {
            int c = 28;
            while (len_base[c] > len)
                --c;
            return c;
        };
        auto dist_code = [&](int dist) {
            int c = 29;
            while (dist_base[c] > dist)
                --c;
            return c;
        };
        // ---- code lengths: fixed, and dynamic (Huffman on the token statistics, limited to 15 bits the way miniz does it)
        std::vector<uint32_t> f_ll(286, 0), f_d(30, 0);
        for (const Token &t : tokens)
            if (t.len) {
                f_ll[257 + len_code(t.len)]++;
                f_d[dist_code(t.dist)]++;
            } else
                f_ll[t.dist]++;
        f_ll[256] = 1;
        auto huff_lengths = [](const std::vector<uint32_t> &freq, int max_bits) {
            const int m = (int)freq.size();
            std::vector<uint8_t> len(m, 0);
            std::vector<int> used;
            for (int i = 0; i < m; ++i)
                if (freq[i])
                    used.push_back(i);
            if (used.empty())
                return len;
            if (used.size() == 1) {
                len[used[0]] = 1;
                return len;
            }
            // plain Huffman tree by repeated merging (m <= 286: quadratic selection is fine)
            struct Node {
                uint64_t w;
                int l, r;
            };
            std::vector<Node> nodes;
            std::vector<int> live;
            for (int i : used) {
                nodes.push_back(Node{freq[i], -1, -1});
                live.push_back((int)nodes.size() - 1);
            }
            while (live.size() > 1) {
                std::sort(live.begin(), live.end(), [&](int a, int b) { return nodes[a].w > nodes[b].w || (nodes[a].w == nodes[b].w && a < b); });
                const int a = live.back();
                live.pop_back();
                const int b = live.back();
                live.pop_back();
                nodes.push_back(Node{nodes[a].w + nodes[b].w, a, b});
                live.push_back((int)nodes.size() - 1);
            }
            std::vector<int> depth(nodes.size(), 0), stack = {live[0]};
            std::vector<int> num(64, 0);
            while (!stack.empty()) {
                const int v = stack.back();
                stack.pop_back();
                if (nodes[v].l < 0) {
                    num[std::min(depth[v], 63)]++;
                    continue;
                }
                depth[nodes[v].l] = depth[nodes[v].r] = depth[v] + 1;
                stack.push_back(nodes[v].l);
                stack.push_back(nodes[v].r);
            }
            // limit to max_bits: fold the deeper leaves into max_bits and repair the Kraft sum (miniz: enforce_max_code_size)
            for (int i = max_bits + 1; i < 64; ++i) {
                num[max_bits] += num[i];
                num[i] = 0;
            }
            uint64_t total = 0;
            for (int i = max_bits; i > 0; --i)
                total += (uint64_t)num[i] << (max_bits - i);
            while (total != (1ull << max_bits)) {
                num[max_bits]--;
                for (int i = max_bits - 1; i > 0; --i)
                    if (num[i]) {
                        num[i]--;
                        num[i + 1] += 2;
                        break;
                    }
                total--;
            }
            // hand the lengths out: most frequent symbols get the shortest codes
            std::sort(used.begin(), used.end(), [&](int a, int b) { return freq[a] > freq[b] || (freq[a] == freq[b] && a < b); });
            size_t k = 0;
            for (int bits = 1; bits <= max_bits; ++bits)
                for (int c = 0; c < num[bits]; ++c)
                    len[used[k++]] = (uint8_t)bits;
            return len;
        };
        auto canonical = [](const std::vector<uint8_t> &len) {
            std::vector<uint16_t> code(len.size(), 0);
            int bl_count[16] = {0}, next[16] = {0};
            for (uint8_t l : len)
                bl_count[l]++;
            bl_count[0] = 0;
            int c = 0;
            for (int b = 1; b < 16; ++b) {
                c = (c + bl_count[b - 1]) << 1;
                next[b] = c;
            }
            for (size_t i = 0; i < len.size(); ++i)
                if (len[i])
                    code[i] = (uint16_t)next[len[i]]++;
            return code;
        };
        std::vector<uint8_t> fix_ll(288), fix_d(30, 5);
        for (int i = 0; i < 288; ++i)
            fix_ll[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        std::vector<uint8_t> dyn_ll = huff_lengths(f_ll, 15), dyn_d = huff_lengths(f_d, 15);
        if (std::count(dyn_d.begin(), dyn_d.end(), 0) == (long)dyn_d.size())
            dyn_d[0] = 1;  // at least one distance code must be described
        auto body_bits = [&](const std::vector<uint8_t> &ll, const std::vector<uint8_t> &d) {
            uint64_t bits = ll[256];
            for (int i = 0; i < 286; ++i)
                bits += (uint64_t)f_ll[i] * ll[i] + (i >= 257 ? (uint64_t)f_ll[i] * len_extra[i - 257] : 0);
            bits -= ll[256];  // (f_ll[256] == 1 is already in the loop)
            for (int i = 0; i < 30; ++i)
                bits += (uint64_t)f_d[i] * (d[i] + dist_extra[i]);
            return bits;
        };
        // dynamic header, code lengths sent without run-length symbols: 5 + 5 + 4 + 19 * 3 + (286 + 30) code-length codes
        std::vector<uint32_t> f_cl(19, 0);
        for (int i = 0; i < 286; ++i)
            f_cl[dyn_ll[i]]++;
        for (int i = 0; i < 30; ++i)
            f_cl[dyn_d[i]]++;
        const std::vector<uint8_t> cl_len = huff_lengths(f_cl, 7);
        uint64_t hdr_bits = 14 + 19 * 3;
        for (int i = 0; i < 19; ++i)
            hdr_bits += (uint64_t)f_cl[i] * cl_len[i];
        const bool use_dynamic = hdr_bits + body_bits(dyn_ll, dyn_d) < body_bits(fix_ll, fix_d);
        const std::vector<uint8_t> &ll_len = use_dynamic ? dyn_ll : fix_ll, &d_len = use_dynamic ? dyn_d : fix_d;
        const std::vector<uint16_t> ll_code = canonical(ll_len), d_code = canonical(d_len), cl_code = canonical(cl_len);
        // ---- bit stream
        uint64_t acc = 0;
        int nbits = 0;
        auto put = [&](uint32_t v, int nb) {  // nb bits, LSB first
            acc |= (uint64_t)v << nbits;
            nbits += nb;
            while (nbits >= 8) {
                z.push_back((uint8_t)(acc & 255));
                acc >>= 8;
                nbits -= 8;
            }
        };
        auto put_code = [&](uint32_t code, int nb) {  // Huffman codes go out MSB first
            uint32_t r = 0;
            for (int i = 0; i < nb; ++i)
                r |= ((code >> i) & 1u) << (nb - 1 - i);
            put(r, nb);
        };
        put(1, 1);                     // BFINAL
        put(use_dynamic ? 2 : 1, 2);   // BTYPE
        if (use_dynamic) {
            static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            put(286 - 257, 5);
            put(30 - 1, 5);
            put(19 - 4, 4);
            for (int i = 0; i < 19; ++i)
                put(cl_len[order[i]], 3);
            for (int i = 0; i < 286; ++i)
                put_code(cl_code[dyn_ll[i]], cl_len[dyn_ll[i]]);
            for (int i = 0; i < 30; ++i)
                put_code(cl_code[dyn_d[i]], cl_len[dyn_d[i]]);
        }
        for (const Token &t : tokens) {
            if (!t.len) {
                put_code(ll_code[t.dist], ll_len[t.dist]);
                continue;
            }
            const int lc = len_code(t.len), dc = dist_code(t.dist);
            put_code(ll_code[257 + lc], ll_len[257 + lc]);
            put((uint32_t)(t.len - len_base[lc]), len_extra[lc]);
            put_code(d_code[dc], d_len[dc]);
            put((uint32_t)(t.dist - dist_base[dc]), dist_extra[dc]);
        }
        put_code(ll_code[256], ll_len[256]);  // end of block
        if (nbits > 0)
            put(0, 8 - nbits);
    }
    put_be32(z, adler32_of(raw.data(), raw.size()));
    chunk("IDAT", z);
    chunk("IEND", {});
}

// ------------------------------------------------------------------ QDataStream (Qt_5_0) on a byte buffer

namespace {

struct Stream {
    std::vector<uint8_t> d;
    size_t pos = 0;
    bool ok = true;
    uint32_t u32()
    {
        if (pos + 4 > d.size()) {
            ok = false;
            return 0;
        }
        const uint32_t v = be32(&d[pos]);
        pos += 4;
        return v;
    }
    bool boolean()
    {
        if (pos + 1 > d.size()) {
            ok = false;
            return false;
        }
        return d[pos++] != 0;
    }
    std::vector<uint8_t> bytes()
    {
        const uint32_t n = u32();
        if (!ok || n == 0xFFFFFFFFu)
            return {};
        if (pos + n > d.size()) {
            ok = false;
            return {};
        }
        std::vector<uint8_t> v(d.begin() + pos, d.begin() + pos + n);
        pos += n;
        return v;
    }
    std::string qstring()  // UTF-16BE -> UTF-8
    {
        const std::vector<uint8_t> b = bytes();
        std::string s;
        for (size_t i = 0; i + 1 < b.size(); i += 2) {
            uint32_t cp = (uint32_t)(b[i] << 8) | b[i + 1];
            if (cp >= 0xD800 && cp < 0xDC00 && i + 3 < b.size()) {
                const uint32_t lo = (uint32_t)(b[i + 2] << 8) | b[i + 3];
                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                i += 2;
            }
            if (cp < 0x80) s += (char)cp;
            else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
            else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
            else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 63)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
        }
        return s;
    }
    // CustomQDataStream::operator>>(cv::Mat&) (CustomQDataStream.h:56-87)
    bool mat(bool png, Image8 &img, std::string &err)
    {
        if (png) {
            const std::vector<uint8_t> b = bytes();
            return ok && png_decode(b.data(), b.size(), img, err);
        }
        const uint32_t type = u32(), rows = u32(), cols = u32();
        const std::vector<uint8_t> b = bytes();
        const uint32_t cn = ((type >> 3) & 511) + 1;
        if (!ok || (type & 7) != 0 || rows > kMaxSide || cols > kMaxSide || b.size() != (size_t)rows * cols * cn) {
            err = "raw image in stream is not 8-bit or has the wrong size";
            return false;
        }
        img.rows = (int)rows;
        img.cols = (int)cols;
        img.channels = (int)cn;
        img.px = b;
        return true;
    }
};

struct Sink {
    std::vector<uint8_t> d;
    void u32(uint32_t v) { put_be32(d, v); }
    void boolean(bool v) { d.push_back(v ? 1 : 0); }
    void bytes(const std::vector<uint8_t> &b)
    {
        u32((uint32_t)b.size());
        d.insert(d.end(), b.begin(), b.end());
    }
    void qstring(const std::string &s)  // UTF-8 -> UTF-16BE
    {
        std::vector<uint8_t> b;
        for (size_t i = 0; i < s.size();) {
            const unsigned char c = (unsigned char)s[i];
            uint32_t cp;
            int n;
            if (c < 0x80) { cp = c; n = 1; }
            else if ((c >> 5) == 6) { cp = c & 31; n = 2; }
            else if ((c >> 4) == 14) { cp = c & 15; n = 3; }
            else { cp = c & 7; n = 4; }
            for (int k = 1; k < n && i + k < s.size(); ++k)
                cp = (cp << 6) | ((unsigned char)s[i + k] & 63);
            i += n;
            auto unit = [&](uint32_t u) { b.push_back((uint8_t)(u >> 8)); b.push_back((uint8_t)(u & 255)); };
            if (cp >= 0x10000) {
                cp -= 0x10000;
                unit(0xD800 + (cp >> 10));
                unit(0xDC00 + (cp & 0x3FF));
            } else
                unit(cp);
        }
        bytes(b);
    }
    void mat_png(const Image8 &img)
    {
        std::vector<uint8_t> b;
        png_encode(img, b);
        bytes(b);
    }
};

bool read_file(const char *path, std::vector<uint8_t> &out, std::string &err)
{
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) {
        err = std::string("File is not readable: ") + (path ? path : "(null)");
        return false;
    }
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0)
        out.insert(out.end(), buf, buf + n);
    fclose(f);
    return true;
}

bool write_file(const char *path, const std::vector<uint8_t> &d, std::string &err)
{
    FILE *f = path ? fopen(path, "wb") : nullptr;
    if (!f) {
        err = std::string("File is not writable: ") + (path ? path : "(null)");
        return false;
    }
    const bool ok = fwrite(d.data(), 1, d.size(), f) == d.size();
    fclose(f);
    if (!ok)
        err = "short write";
    return ok;
}

constexpr uint32_t kMcsMagic = 0x87AECFB1u, kMcsVersion = 8, kMcsVersionEncoded = 8;  // CellShape.h:31-36
constexpr uint32_t kMilMagic = 0xADBE2480u, kMilVersion = 6, kMilVersionEncoded = 6;  // ImageLibrary.h:12-18

}  // namespace

// ------------------------------------------------------------------ .mcs

bool load_mcs(const char *path, McsFile &out, std::string &err)
{
    Stream s;
    if (!read_file(path, s.d, err))
        return false;
    if (s.u32() != kMcsMagic || !s.ok) {
        err = "File is not a valid .mcs";  // CellShape.cpp:381-383
        return false;
    }
    const uint32_t version = s.u32();
    if (!(version <= kMcsVersion && version >= 7)) {  // CellShape.cpp:388-398
        err = version < kMcsVersion ? ".mcs uses an outdated file version" : ".mcs uses a newer file version";
        return false;
    }
    out = McsFile();
    out.version = version;
    out.name = s.qstring();
    Image8 img;
    if (!s.mat(version >= kMcsVersionEncoded, img, err))
        return false;
    if (img.rows != img.cols || img.rows <= 0) {
        err = "cell mask in the .mcs is not square";
        return false;
    }
    Shape &sh = out.shape;
    sh.size = img.rows;
    sh.mask.resize((size_t)img.rows * img.cols);
    for (size_t i = 0; i < sh.mask.size(); ++i)
        sh.mask[i] = img.px[i * img.channels];  // stored as decoded: loadFromFile does not threshold (CellShape.cpp:405-410)
    int32_t v[6];
    for (int i = 0; i < 6; ++i)
        v[i] = (int32_t)s.u32();
    sh.row_spacing = v[0]; sh.col_spacing = v[1];
    sh.alt_row_spacing = v[2]; sh.alt_col_spacing = v[3];
    sh.alt_row_offset = v[4]; sh.alt_col_offset = v[5];
    sh.alt_col_flip_h = s.boolean(); sh.alt_col_flip_v = s.boolean();
    sh.alt_row_flip_h = s.boolean(); sh.alt_row_flip_v = s.boolean();
    if (!s.ok) {
        err = "truncated .mcs";
        return false;
    }
    return true;
}

bool save_mcs(const char *path, const McsFile &in, std::string &err)
{
    const Shape &sh = in.shape;
    if (sh.size <= 0 || sh.mask.size() != (size_t)sh.size * sh.size) {
        err = "cell shape has no mask";
        return false;
    }
    Sink k;
    k.u32(kMcsMagic);
    k.u32(kMcsVersion);
    k.qstring(in.name);
    Image8 img;
    img.rows = img.cols = sh.size;
    img.channels = 1;
    img.px = sh.mask;
    k.mat_png(img);
    const int v[6] = {sh.row_spacing, sh.col_spacing, sh.alt_row_spacing, sh.alt_col_spacing, sh.alt_row_offset, sh.alt_col_offset};
    for (int x : v)
        k.u32((uint32_t)x);
    k.boolean(sh.alt_col_flip_h); k.boolean(sh.alt_col_flip_v);
    k.boolean(sh.alt_row_flip_h); k.boolean(sh.alt_row_flip_v);
    return write_file(path, k.d, err);
}

// ------------------------------------------------------------------ .mil

bool load_mil(const char *path, MilFile &out, std::string &err)
{
    Stream s;
    if (!read_file(path, s.d, err))
        return false;
    if (s.u32() != kMilMagic || !s.ok) {
        err = "File is not a valid .mil";  // ImageLibrary.cpp:174-176
        return false;
    }
    const uint32_t version = s.u32();
    if (!(version <= kMilVersion && version >= 4)) {  // ImageLibrary.cpp:181-191
        err = version < kMilVersion ? ".mil uses an outdated file version" : ".mil uses a newer file version";
        return false;
    }
    out = MilFile();
    out.version = version;
    out.image_size = (int)s.u32();
    const uint32_t n = s.u32();
    if (!s.ok || out.image_size < 0) {
        err = "truncated .mil";
        return false;
    }
    const size_t per = (size_t)out.image_size * out.image_size * 3;
    for (uint32_t i = 0; i < n; ++i) {
        Image8 img;
        if (!s.mat(version >= kMilVersionEncoded, img, err))
            return false;
        out.names.push_back(s.qstring());
        if (!s.ok) {
            err = "truncated .mil";
            return false;
        }
        if (img.rows != out.image_size || img.cols != out.image_size) {
            err = "image in the .mil does not have the library's image size";
            return false;
        }
        const size_t base = out.images.size();
        out.images.resize(base + per);
        for (size_t p = 0; p < (size_t)img.rows * img.cols; ++p)
            for (int c = 0; c < 3; ++c)  // grey -> 3 equal channels, BGRA -> BGR
                out.images[base + p * 3 + c] = img.px[p * img.channels + (img.channels >= 3 ? c : 0)];
    }
    // versions < 5 are shuffled on load by the reference (std::random_device, ImageLibrary.cpp:219-228): file order is kept here
    return true;
}

bool save_mil(const char *path, const MilFile &in, std::string &err)
{
    const size_t per = (size_t)in.image_size * in.image_size * 3;
    if (in.image_size <= 0 ? !in.images.empty() : in.images.size() % per != 0) {
        err = "library images do not match the image size";
        return false;
    }
    const size_t n = per ? in.images.size() / per : 0;
    Sink k;
    k.u32(kMilMagic);
    k.u32(kMilVersion);
    k.u32((uint32_t)in.image_size);
    k.u32((uint32_t)n);
    for (size_t i = 0; i < n; ++i) {
        Image8 img;
        img.rows = img.cols = in.image_size;
        img.channels = 3;
        img.px.assign(in.images.begin() + i * per, in.images.begin() + (i + 1) * per);
        k.mat_png(img);
        k.qstring(i < in.names.size() ? in.names[i] : std::string());
    }
    return write_file(path, k.d, err);
}

}  // namespace mm
