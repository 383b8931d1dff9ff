#include "opr.h"
#include <cstdio>
#include <vector>
int main(int argc, char** argv) {
    int w = 0, h = 0;
    if (opr_image_info(argv[1], &w, &h)) { printf("err %.80s\n", opr_io_last_error()); return 0; }
    if (w <= 0 || h <= 0 || (long long)w * h > (1 << 26)) { printf("big %d %d\n", w, h); return 0; }
    std::vector<unsigned char> px((size_t)w * h * 4);
    if (opr_image_read(argv[1], px.data(), px.size())) { printf("err2 %.80s\n", opr_io_last_error()); return 0; }
    unsigned long acc = 0; for (unsigned char c : px) acc += c;
    printf("ok %d %d %lu\n", w, h, acc);
    // a too-small destination must be refused, not overrun
    std::vector<unsigned char> small(16);
    int rc = opr_image_read(argv[1], small.data(), small.size());
    printf("small rc=%d\n", rc);
    return 0;
}
