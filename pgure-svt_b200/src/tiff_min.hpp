// Minimal multi-page baseline TIFF reader/writer for the CLI (the reference uses libtiff, whose headers are not
// available here: src/PGURE-SVT.cpp:103-160,206-239).  Reads uncompressed (Compression = 1) single-sample 8- or
// 16-bit strips, little- or big-endian; writes 16-bit little-endian, one strip and one IFD per page with the tags
// the reference sets (WIDTH, LENGTH, BITSPERSAMPLE=16, SAMPLESPERPIXEL=1, PLANARCONFIG=CONTIG,
// PHOTOMETRIC=MINISBLACK, ORIENTATION=TOPLEFT, SUBFILETYPE=PAGE, PAGENUMBER).
#ifndef PGURESVT_B200_TIFF_MIN_HPP
#define PGURESVT_B200_TIFF_MIN_HPP
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace tiffmin
{
struct Page
{
    uint32_t width = 0, height = 0, rows_per_strip = 0;
    uint16_t bits = 0, compression = 1, samples = 1;
    std::vector<uint32_t> offsets, counts;
};

class Reader
{
public:
    std::string error;
    bool open(const std::string &path)
    {
        f = std::fopen(path.c_str(), "rb");
        if (!f)
            return fail("cannot open " + path);
        unsigned char h[8];
        if (std::fread(h, 1, 8, f) != 8)
            return fail("short header");
        if (h[0] == 'I' && h[1] == 'I')
            big = false;
        else if (h[0] == 'M' && h[1] == 'M')
            big = true;
        else
            return fail("not a TIFF file");
        if (rd16(h + 2) != 42)
            return fail("not a classic TIFF (BigTIFF is not supported)");
        uint32_t ifd = rd32(h + 4);
        while (ifd)
        {
            Page pg;
            if (!read_ifd(ifd, pg, ifd))
                return false;
            pages.push_back(pg);
        }
        return !pages.empty() || fail("no image directories");
    }
    ~Reader()
    {
        if (f)
            std::fclose(f);
    }
    size_t n_pages() const { return pages.size(); }
    const Page &page(size_t i) const { return pages[i]; }
    // row-major scanlines widened to uint16 (the reference reads 8-bit scanlines into a uint16 buffer without
    // widening, which scrambles them — SURVEY §8f rank 1; here 8-bit input is widened properly)
    bool read_page(size_t i, uint16_t *out)
    {
        const Page &pg = pages[i];
        if (pg.compression != 1)
            return fail("compressed TIFF pages are not supported");
        if (pg.samples != 1 || (pg.bits != 8 && pg.bits != 16))
            return fail("only single-sample 8/16-bit pages are supported");
        const size_t bps = pg.bits / 8, total = (size_t)pg.width * pg.height * bps;
        std::vector<unsigned char> raw(total);
        size_t pos = 0;
        for (size_t s = 0; s < pg.offsets.size() && pos < total; s++)
        {
            size_t n = pg.counts.size() > s ? pg.counts[s] : 0;
            if (n == 0 || pos + n > total)
                n = total - pos;
            if (std::fseek(f, (long)pg.offsets[s], SEEK_SET) != 0 || std::fread(raw.data() + pos, 1, n, f) != n)
                return fail("truncated strip");
            pos += n;
        }
        const size_t npx = (size_t)pg.width * pg.height;
        if (bps == 1)
            for (size_t k = 0; k < npx; k++)
                out[k] = raw[k];
        else
            for (size_t k = 0; k < npx; k++)
                out[k] = big ? (uint16_t)((raw[2 * k] << 8) | raw[2 * k + 1]) : (uint16_t)(raw[2 * k] | (raw[2 * k + 1] << 8));
        return true;
    }

private:
    FILE *f = nullptr;
    bool big = false;
    std::vector<Page> pages;
    bool fail(const std::string &m)
    {
        error = m;
        return false;
    }
    uint16_t rd16(const unsigned char *p) const { return big ? (uint16_t)((p[0] << 8) | p[1]) : (uint16_t)(p[0] | (p[1] << 8)); }
    uint32_t rd32(const unsigned char *p) const
    {
        return big ? ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]
                   : ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0];
    }
    bool values(uint16_t type, uint32_t count, const unsigned char *field, std::vector<uint32_t> &out)
    {
        const size_t sz = (type == 3) ? 2 : (type == 4) ? 4 : (type == 1) ? 1 : 0;
        if (!sz)
            return fail("unsupported tag type");
        if (count == 0 || count > (1u << 26)) // (the callers index out[0]; a hostile count must not size a buffer)
            return fail("tag with an empty or absurd value count");
        std::vector<unsigned char> buf((size_t)count * sz);
        if (buf.size() <= 4)
            std::memcpy(buf.data(), field, buf.size());
        else
        {
            const long keep = std::ftell(f);
            if (std::fseek(f, (long)rd32(field), SEEK_SET) != 0 || std::fread(buf.data(), 1, buf.size(), f) != buf.size())
                return fail("truncated tag data");
            std::fseek(f, keep, SEEK_SET);
        }
        out.resize(count);
        for (uint32_t k = 0; k < count; k++)
            out[k] = sz == 2 ? rd16(&buf[2 * k]) : sz == 4 ? rd32(&buf[4 * k]) : buf[k];
        return true;
    }
    bool read_ifd(uint32_t off, Page &pg, uint32_t &next)
    {
        unsigned char b[12];
        if (std::fseek(f, (long)off, SEEK_SET) != 0 || std::fread(b, 1, 2, f) != 2)
            return fail("bad IFD offset");
        const uint16_t n = rd16(b);
        for (uint16_t e = 0; e < n; e++)
        {
            if (std::fread(b, 1, 12, f) != 12)
                return fail("truncated IFD");
            const uint16_t tag = rd16(b), type = rd16(b + 2);
            const uint32_t count = rd32(b + 4);
            std::vector<uint32_t> v;
            switch (tag)
            {
            case 256: if (!values(type, count, b + 8, v)) return false; pg.width = v[0]; break;
            case 257: if (!values(type, count, b + 8, v)) return false; pg.height = v[0]; break;
            case 258: if (!values(type, count, b + 8, v)) return false; pg.bits = (uint16_t)v[0]; break;
            case 259: if (!values(type, count, b + 8, v)) return false; pg.compression = (uint16_t)v[0]; break;
            case 273: if (!values(type, count, b + 8, pg.offsets)) return false; break;
            case 277: if (!values(type, count, b + 8, v)) return false; pg.samples = (uint16_t)v[0]; break;
            case 278: if (!values(type, count, b + 8, v)) return false; pg.rows_per_strip = v[0]; break;
            case 279: if (!values(type, count, b + 8, pg.counts)) return false; break;
            default: break;
            }
        }
        if (std::fread(b, 1, 4, f) != 4)
            return fail("truncated IFD");
        next = rd32(b);
        if (pg.bits == 0)
            pg.bits = 1;
        return true;
    }
};

class Writer
{
public:
    bool open(const std::string &path)
    {
        f = std::fopen(path.c_str(), "wb");
        if (!f)
            return false;
        const unsigned char h[8] = {'I', 'I', 42, 0, 0, 0, 0, 0};
        std::fwrite(h, 1, 8, f);
        link_pos = 4;
        return true;
    }
    // row-major scanlines, 16-bit
    bool write_page(const uint16_t *px, uint32_t width, uint32_t height, uint16_t page, uint16_t n_pages)
    {
        long pos = std::ftell(f);
        if (pos & 1)
        {
            std::fputc(0, f);
            pos++;
        }
        const uint32_t data_off = (uint32_t)pos, nbytes = width * height * 2;
        std::fwrite(px, 1, nbytes, f);
        const uint32_t ifd_off = data_off + nbytes;
        struct E
        {
            uint16_t tag, type;
            uint32_t count, value;
        };
        const E entries[] = {{254, 4, 1, 2}, {256, 4, 1, width}, {257, 4, 1, height}, {258, 3, 1, 16}, {259, 3, 1, 1},
                             {262, 3, 1, 1}, {273, 4, 1, data_off}, {274, 3, 1, 1}, {277, 3, 1, 1}, {278, 4, 1, height},
                             {279, 4, 1, nbytes}, {284, 3, 1, 1}, {297, 3, 2, (uint32_t)page | ((uint32_t)n_pages << 16)}};
        const uint16_t n = sizeof(entries) / sizeof(entries[0]);
        put16(n);
        for (const E &e : entries)
        {
            put16(e.tag);
            put16(e.type);
            put32(e.count);
            put32(e.value);
        }
        const long next_pos = std::ftell(f);
        put32(0);
        // patch the previous link
        std::fseek(f, link_pos, SEEK_SET);
        put32(ifd_off);
        std::fseek(f, 0, SEEK_END);
        link_pos = next_pos;
        return !std::ferror(f);
    }
    void close()
    {
        if (f)
            std::fclose(f);
        f = nullptr;
    }
    ~Writer() { close(); }

private:
    FILE *f = nullptr;
    long link_pos = 4;
    void put16(uint16_t v)
    {
        const unsigned char b[2] = {(unsigned char)(v & 255), (unsigned char)(v >> 8)};
        std::fwrite(b, 1, 2, f);
    }
    void put32(uint32_t v)
    {
        const unsigned char b[4] = {(unsigned char)(v & 255), (unsigned char)((v >> 8) & 255), (unsigned char)((v >> 16) & 255),
                                    (unsigned char)(v >> 24)};
        std::fwrite(b, 1, 4, f);
    }
};
} // namespace tiffmin
#endif
