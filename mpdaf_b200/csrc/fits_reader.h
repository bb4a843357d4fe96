// Minimal FITS image-HDU access for the file-name entry points.
//
// Replaces what the reference gets from cfitsio inside src/merging.c:93-110
// (open "file[EXTNAME]", require a 3-D image, report its shape) and
// :228/:411/:417 (read runs of pixels).  Only what the path needs: primary HDU +
// IMAGE extensions, BITPIX -32 / -64, no scaling keywords, no compression.
// Pixels are returned as raw big-endian bytes; the byte swap happens on the GPU.
#pragma once
#include <string>

namespace mbfits {

struct ImageHdu {
    int fd = -1;
    long long data_offset = 0;  // byte offset of the first pixel
    int bitpix = 0;             // -32 or -64
    int naxis = 0;
    long long naxes[3] = {1, 1, 1};  // x (fastest), y, z
    std::string path;
};

// Opens `path`, walks the HDU list until EXTNAME == extname (case-insensitive).
// Returns 0 or an MPDAF_B200_ERR_* code; `err` receives a message.
int open_image(const std::string& path, const std::string& extname, ImageHdu* out, std::string* err);
void close_image(ImageHdu* h);

// Reads `count` pixels starting at flat pixel index `first` as raw file bytes
// (big-endian) into `dst`.  Thread-safe (pread).
int read_raw(const ImageHdu& h, long long first, long long count, void* dst, std::string* err);

}  // namespace mbfits
