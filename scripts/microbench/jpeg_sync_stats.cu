// Offline statistics of the self-synchronising JPEG decode (host build of csrc/jpeg_decode.cu, no GPU): per round, how many
// subsequences are decoded again / change / still end in a wrong state, what is wrong about them after round 0 (bit position,
// coefficient index, block-in-MCU phase), and the same for a variant that starts every thread one or two subsequences early.
//   nvcc -O2 -std=c++17 --expt-relaxed-constexpr -w -o jpeg_sync_stats jpeg_sync_stats.cu ; ./jpeg_sync_stats <request file> [subsequence bytes]
// (request file = the one tests/test_jpeg_host.py::run_host_tool writes: int32 header[32], table set, entropy-coded data)
#define FB_JPEG_HOST_TEST 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../facet_b200/csrc/jpeg_decode.cu"
namespace fb { void set_error(const char*, ...) {} const char* get_error() { return ""; } int sm_count() { return 1; } }
using namespace fb;
static int wrongstat[2];
static bool eq(const SyncState& a, const SyncState& b) { return a.pos == b.pos && a.b == b.b && a.k == b.k; }
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    int32_t hdr[32]; fread(hdr, 4, 32, f);
    JpegTableSet* T_ = (JpegTableSet*)aligned_alloc(16, sizeof(JpegTableSet));
    fread(T_, 1, sizeof(JpegTableSet), f);
    const long long len = hdr[17];
    std::vector<uint8_t> scan(len + 16); fread(scan.data(), 1, len, f); fclose(f);
    JpegGeom g; g.width = hdr[0]; g.height = hdr[1]; g.ncomp = hdr[2];
    for (int c = 0; c < 3; ++c) { g.hs[c] = hdr[3 + c]; g.vs[c] = hdr[6 + c]; g.tq[c] = hdr[9 + c]; g.td[c] = hdr[12 + c]; g.ta[c] = hdr[18 + c]; }
    g.restart_interval = 0;
    g.mcux = (g.width + 8 * g.hs[0] - 1) / (8 * g.hs[0]); g.mcuy = (g.height + 8 * g.vs[0] - 1) / (8 * g.vs[0]);
    std::vector<uint8_t> clean; clean.reserve(len + 64);
    for (long long p = 0; p < len; ++p) if (!(scan[p] == 0x00 && p > 0 && scan[p - 1] == 0xFF)) clean.push_back(scan[p]);
    const long long clean_len = clean.size(), len_bits = 8 * clean_len; clean.resize(clean_len + 64, 0);
    const int S = argc > 2 ? atoi(argv[2]) : kSubseqBytes;
    const int T = (int)((clean_len + S - 1) / S);
    int nblk; const uint64_t lay = mcu_layout(g, nblk);
    // truth
    std::vector<SyncState> truth(T);
    { SyncState st; st.pos = 0; st.b = st.k = 0;
      for (int t = 0; t < T; ++t) { SyncState out; int done; span_decode<false>(clean.data(), len_bits, st, (long long)(t + 1) * S * 8, g, *T_, h_zigzag, lay, nblk, out, done, 0, 0, nullptr); truth[t] = out; st = out; } }
    // scheme A
    std::vector<SyncState> cur(T), prev(T), prev2(T);
    printf("T=%d subseq=%d B\nscheme A (guess at own start):\n", T, S);
    for (int round = 0; round <= 30; ++round) {
        int active = 0, changed = 0, wrong = 0;
        for (int t = 0; t < T; ++t) {
            SyncState st; const long long start = (long long)t * S * 8;
            if (round == 0 || t == 0) { st.pos = start; st.b = st.k = 0; } else st = prev[t - 1];
            if (round >= 2 && t >= 1 && eq(prev[t - 1], prev2[t - 1])) { cur[t] = prev[t]; }
            else { SyncState out; int done; span_decode<false>(clean.data(), len_bits, st, start + (long long)S * 8, g, *T_, h_zigzag, lay, nblk, out, done, 0, 0, nullptr); cur[t] = out; ++active; }
            if (round >= 1 && !eq(cur[t], prev[t])) ++changed;
            if (!eq(cur[t], truth[t])) { ++wrong; if (round == 0) { static int c_pos=0,c_posk=0; if (cur[t].pos==truth[t].pos) { ++c_pos; if (cur[t].k==truth[t].k) ++c_posk; } if (t==T-1|| true) { wrongstat[0]=c_pos; wrongstat[1]=c_posk; } } }
        }
        printf("  round %2d: decoded %6d  changed %6d  wrong end states %6d\n", round, active, changed, wrong);
        if (round == 0) printf("     of the wrong ones: pos right %d, pos and k right (only the block phase wrong) %d\n", wrongstat[0], wrongstat[1]);
        prev2 = prev; prev = cur;
        if (round >= 1 && !changed) break;
    }
    // scheme B: start one subsequence earlier
    for (int back = 1; back <= 2; ++back) {
      std::vector<SyncState> mid(T), end(T);
      int wrong_mid = 0, wrong_end = 0;
      for (int t = 0; t < T; ++t) {
        SyncState st; const int t0 = t - back < 0 ? 0 : t - back;
        st.pos = (long long)t0 * S * 8; st.b = st.k = 0;
        SyncState out; int done;
        if (t0 < t) { span_decode<false>(clean.data(), len_bits, st, (long long)t * S * 8, g, *T_, h_zigzag, lay, nblk, out, done, 0, 0, nullptr); st = out; }
        mid[t] = st;
        span_decode<false>(clean.data(), len_bits, st, (long long)(t + 1) * S * 8, g, *T_, h_zigzag, lay, nblk, out, done, 0, 0, nullptr);
        end[t] = out;
        if (t > 0 && !eq(mid[t], truth[t - 1])) ++wrong_mid;
        if (!eq(end[t], truth[t])) ++wrong_end;
      }
      // rounds: thread t active if mid[t] != end[t-1]
      printf("scheme B (start %d subsequence(s) earlier): wrong start states %d, wrong end states %d\n", back, wrong_mid, wrong_end);
      std::vector<SyncState> e = end, m = mid;
      for (int round = 1; round <= 30; ++round) {
        int active = 0; std::vector<SyncState> ne = e;
        for (int t = 1; t < T; ++t) {
            if (!eq(m[t], e[t - 1])) { SyncState out; int done; span_decode<false>(clean.data(), len_bits, e[t - 1], (long long)(t + 1) * S * 8, g, *T_, h_zigzag, lay, nblk, out, done, 0, 0, nullptr); ne[t] = out; m[t] = e[t - 1]; ++active; }
        }
        e = ne;
        printf("  round %2d: decoded %6d\n", round, active);
        if (!active) break;
      }
      int bad = 0; for (int t = 0; t < T; ++t) if (!eq(e[t], truth[t])) ++bad;
      printf("  final wrong: %d\n", bad);
    }
    return 0;
}
