// See fits_reader.h.  FITS structure per the FITS 4.0 standard: 2880-byte blocks,
// 80-character header cards, header ends with an END card, data padded to a
// block boundary; data size = |BITPIX|/8 * GCOUNT * (PCOUNT + prod NAXISn).
#define _FILE_OFFSET_BITS 64
#include "fits_reader.h"

#include <cctype>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/resource.h>
#include <unistd.h>

#include "../../include/mpdaf_b200.h"

namespace mbfits {
namespace {

constexpr int kBlock = 2880;
constexpr int kCard = 80;

std::string keyword_of(const char* card)
{
    int n = 8;
    while (n > 0 && card[n - 1] == ' ') n--;
    return std::string(card, n);
}

bool has_value(const char* card) { return card[8] == '=' && card[9] == ' '; }

long long int_value(const char* card) { return std::atoll(std::string(card + 10, 70).c_str()); }
double real_value(const char* card) { return std::atof(std::string(card + 10, 70).c_str()); }

std::string string_value(const char* card)
{
    const char* end = card + kCard;
    const char* p = card + 10;
    while (p < end && *p != '\'') p++;
    if (p >= end) return "";
    std::string s;
    for (p++; p < end; p++) {
        if (*p == '\'') {
            if (p + 1 < end && p[1] == '\'') { s.push_back('\''); p++; continue; }
            break;
        }
        s.push_back(*p);
    }
    while (!s.empty() && s.back() == ' ') s.pop_back();
    return s;
}

bool iequal(const std::string& a, const std::string& b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++)
        if (std::toupper((unsigned char)a[i]) != std::toupper((unsigned char)b[i])) return false;
    return true;
}

}  // namespace

int open_image(const std::string& path, const std::string& extname, ImageHdu* out, std::string* err)
{
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0 && errno == EMFILE) {
        // 500 files x (DATA + STAT) is more than the usual soft limit of 1024 descriptors: raise it to the hard limit once
        struct rlimit rl;
        if (getrlimit(RLIMIT_NOFILE, &rl) == 0 && rl.rlim_cur < rl.rlim_max) {
            rl.rlim_cur = rl.rlim_max;
            if (setrlimit(RLIMIT_NOFILE, &rl) == 0) fd = ::open(path.c_str(), O_RDONLY);
        }
    }
    if (fd < 0) {
        *err = (errno == EMFILE ? "too many open files (RLIMIT_NOFILE) opening " : "cannot open ") + path;
        return MPDAF_B200_ERR_IO;
    }
    long long off = 0;
    char blk[kBlock];
    for (int hdu = 0;; hdu++) {
        int bitpix = 0, naxis = 0;
        long long ax[8] = {0}, pcount = 0, gcount = 1;
        double bscale = 1.0, bzero = 0.0;
        std::string name;
        bool done = false;
        while (!done) {
            if (::pread(fd, blk, kBlock, off) != kBlock) {
                ::close(fd);
                *err = "no image extension '" + extname + "' in " + path;
                return MPDAF_B200_ERR_IO;
            }
            off += kBlock;
            for (int c = 0; c < kBlock / kCard && !done; c++) {
                const char* card = blk + c * kCard;
                const std::string key = keyword_of(card);
                if (key == "END") done = true;
                else if (!has_value(card)) continue;
                else if (key == "BITPIX") bitpix = (int)int_value(card);
                else if (key == "NAXIS") naxis = (int)int_value(card);
                else if (key == "PCOUNT") pcount = int_value(card);
                else if (key == "GCOUNT") gcount = int_value(card);
                else if (key == "EXTNAME") name = string_value(card);
                else if (key == "BSCALE") bscale = real_value(card);
                else if (key == "BZERO") bzero = real_value(card);
                else if (key.size() == 6 && key.compare(0, 5, "NAXIS") == 0 && key[5] >= '1' && key[5] <= '8')
                    ax[key[5] - '1'] = int_value(card);   // NAXIS1 .. NAXIS8 only: a NAXIS0 / NAXIS9 card must not index ax[]
            }
        }
        if (naxis < 0 || naxis > 8) {
            ::close(fd);
            *err = path + ": NAXIS out of range";
            return MPDAF_B200_ERR_IO;
        }
        long long nelem = naxis > 0 ? 1 : 0;
        for (int k = 0; k < naxis; k++) nelem *= ax[k];
        const long long bytes = (std::llabs((long long)bitpix) / 8) * gcount * (pcount + nelem);
        if ((extname.empty() && hdu == 0) || (!extname.empty() && iequal(name, extname))) {
            if (naxis != 3) {  // merging.c:104-107
                ::close(fd);
                *err = path + "[" + extname + "] is not a cube";
                return MPDAF_B200_ERR_SHAPE;
            }
            if (bitpix != -32 && bitpix != -64) {
                ::close(fd);
                *err = path + "[" + extname + "]: only BITPIX -32/-64 images are supported";
                return MPDAF_B200_ERR_IO;
            }
            if (bscale != 1.0 || bzero != 0.0) {   // cfitsio would apply them while reading (merging.c:228); this reader does not
                ::close(fd);
                *err = path + "[" + extname + "]: scaled images (BSCALE / BZERO) are not supported";
                return MPDAF_B200_ERR_IO;
            }
            if (ax[0] < 0 || ax[1] < 0 || ax[2] < 0) {
                ::close(fd);
                *err = path + "[" + extname + "]: negative axis length";
                return MPDAF_B200_ERR_IO;
            }
            out->fd = fd;
            out->data_offset = off;
            out->bitpix = bitpix;
            out->naxis = naxis;
            for (int k = 0; k < 3; k++) out->naxes[k] = ax[k];
            out->path = path;
            return MPDAF_B200_OK;
        }
        off += (bytes + kBlock - 1) / kBlock * kBlock;
    }
}

void close_image(ImageHdu* h)
{
    if (h->fd >= 0) ::close(h->fd);
    h->fd = -1;
}

int read_raw(const ImageHdu& h, long long first, long long count, void* dst, std::string* err)
{
    const int esz = -h.bitpix / 8;
    char* p = static_cast<char*>(dst);
    long long todo = count * esz;
    long long pos = h.data_offset + first * esz;
    while (todo > 0) {
        const ssize_t got = ::pread(h.fd, p, (size_t)todo, pos);
        if (got <= 0) {
            *err = "short read from " + h.path;
            return MPDAF_B200_ERR_IO;
        }
        p += got;
        pos += got;
        todo -= got;
    }
    return MPDAF_B200_OK;
}

}  // namespace mbfits
