// Host build of the device's DEFLATE decoder (nextpolish2_b200/csrc/np2_inflate.cuh is one source for both): lets the
// non-GPU tests compare it with zlib member by member.  The payload is copied to every 4-byte misalignment, with the
// slack the decoder's word reads need.
#include <cstring>
#include <vector>

#include "np2_inflate.cuh"
extern "C" int np2t_inflate(const uint8_t *payload, uint32_t clen, uint8_t *out, uint32_t cap, uint32_t *produced) {
    int ok = 1;
    for (int mis = 0; mis < 4; mis++) {
        std::vector<uint8_t> buf((size_t)clen + 64, (uint8_t)(0xA5 + mis));
        if (clen) memcpy(buf.data() + 16 + mis, payload, clen);
        uint32_t n = 0;
        const int r = np2::infl::inflate_member_host(buf.data() + 16 + mis, clen, out, cap, &n) ? 1 : 0;
        if (mis == 0) {
            ok = r;
            *produced = n;
        } else if (r != ok || (r && n != *produced)) {
            return -1;  // the result must not depend on the alignment
        }
    }
    return ok;
}
