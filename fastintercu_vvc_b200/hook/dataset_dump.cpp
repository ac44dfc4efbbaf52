// dataset_dump.cpp -- see dataset_dump.h.
#include "dataset_dump.h"

#include <sys/stat.h>

#include <vector>

namespace mlt_hook {

namespace {

uint32_t crc32_update(uint32_t crc, const uint8_t *p, size_t n)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return crc;
}

void put_be32(std::vector<uint8_t> &v, uint32_t x)
{
    v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x));
}

void chunk(std::vector<uint8_t> &out, const char type[4], const std::vector<uint8_t> &data)
{
    put_be32(out, (uint32_t)data.size());
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put_be32(out, crc32_update(0xFFFFFFFFu, out.data() + start, out.size() - start) ^ 0xFFFFFFFFu);
}

bool make_dir(const std::string &p)
{
    if (mkdir(p.c_str(), 0777) == 0) return true;
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

} // namespace

bool writePng16(const std::string &path, const uint16_t *pix, int width, int height)
{
    if (!pix || width <= 0 || height <= 0) return false;
    std::vector<uint8_t> raw; // filter byte + big-endian samples per scanline
    raw.reserve((size_t)height * (1 + 2 * width));
    for (int y = 0; y < height; y++) {
        raw.push_back(0);
        for (int x = 0; x < width; x++) {
            const uint16_t v = pix[(size_t)y * width + x];
            raw.push_back(uint8_t(v >> 8));
            raw.push_back(uint8_t(v));
        }
    }
    std::vector<uint8_t> z = {0x78, 0x01}; // zlib header, then stored (uncompressed) deflate blocks
    uint32_t a = 1, b = 0;                 // Adler-32 of the raw stream
    for (uint8_t c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size();) {
        const size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + n == raw.size() ? 1 : 0);
        z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8));
        z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        off += n;
    }
    put_be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)width);
    put_be32(ihdr, (uint32_t)height);
    ihdr.push_back(16); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); // 16-bit grey, no interlace
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    return std::fclose(f) == 0 && ok;
}

DatasetDump::DatasetDump(const std::string &root, const std::string &sequence, int baseQP, const std::string &csvPath)
    : m_root(root), m_seq(sequence), m_baseQP(baseQP)
{
    if (!make_dir(root) || !make_dir(root + "/" + sequence) || !make_dir(root + "/" + sequence + "/org") ||
        !make_dir(root + "/" + sequence + "/pred"))
        return;
    m_csv = std::fopen(csvPath.c_str(), "a");
}

DatasetDump::~DatasetDump()
{
    if (m_csv) std::fclose(m_csv);
}

bool DatasetDump::dump(const int16_t *org, int orgStride, const int16_t *pred, int predStride, int cuw, int cuh, int poc, int x, int y,
                       int label, int cuQP)
{
    if (!m_csv || !org || !pred) return false;
    std::vector<uint16_t> o((size_t)cuw * cuh), p((size_t)cuw * cuh);
    for (int i = 0; i < cuh; i++)
        for (int j = 0; j < cuw; j++) { // the hook's (uint16_t) casts, EncCu.cpp:816,827
            o[(size_t)i * cuw + j] = (uint16_t)org[(size_t)i * orgStride + j];
            p[(size_t)i * cuw + j] = (uint16_t)pred[(size_t)i * predStride + j];
        }
    char name[96];
    std::snprintf(name, sizeof name, "%d_%d_%d_%d.png", m_baseQP, poc, x, y); // baseQP_POC_X_Y.png (dataset.py:52)
    const std::string base = m_root + "/" + m_seq;
    if (!writePng16(base + "/org/" + name, o.data(), cuw, cuh) || !writePng16(base + "/pred/" + name, p.data(), cuw, cuh)) return false;
    return std::fprintf(m_csv, "%s,%d,%d,%d,%d,%d,%d\n", m_seq.c_str(), m_baseQP, poc, x, y, label, cuQP) > 0 && std::fflush(m_csv) == 0;
}

} // namespace mlt_hook
