// tests/emu/asan_sketch_driver.cpp -- TEST INFRASTRUCTURE ONLY: drives the emulated sketch kernels (emu_sketch.cpp) over random
// jobs with exact-size heap buffers, to be built with -fsanitize=address: any read or write of the kernels outside the
// sequence buffer, the output, the job / status arrays or the block's shared memory aborts the program.
// (tests/test_emu_logic.py::test_emu_sketch_address_sanitizer builds and runs it.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
extern "C" long emu_sketch_jobs(int njobs, const int64_t *seq_off, const int32_t *len, const int32_t *shift, const uint32_t *rid, const char *buf, int w, int k, const char *Z, int W, int small, int grid, int64_t *out_off, uint64_t *out, int64_t out_cap);
extern "C" int emu_sketch_packed(int njobs, const int64_t *seq_off, const int32_t *len, const int32_t *shift, const uint32_t *rid, const char *buf, int w, int k, const char *Z, int W, int pack, int threads, int grid, int64_t stride, int32_t *out_cnt, uint64_t *out);
extern "C" void emu_sketch_concurrency(unsigned seed);
int main(){
  srand(7);
  const char *pats[] = {"10","110","1","100","101001"};
  int cfg[][2] = {{21,11},{15,10},{19,19},{28,8},{12,5},{17,30},{16,9},{4,33}};
  for (int it = 0; it < 16; ++it) {
    int k = cfg[it%8][0], w = cfg[it%8][1]; const char *Z = pats[it%5]; int W = strlen(Z);
    std::vector<int> lens;
    int small = it & 1;
    int tp = 256 - (2*w+k-3) - (w-1);
    int ones = 0; for (int i=0;i<W;++i) ones += Z[i]=='1';
    for (int j = 0; j < 5; ++j) lens.push_back(small ? W + rand() % (tp*W/ones - W > 1 ? tp*W/ones - W : 1) : 40 + rand() % 9000);
    std::vector<int64_t> off; std::vector<int32_t> len, sh; std::vector<uint32_t> rid; 
    size_t tot = 0; for (int l : lens) tot += l;
    // exact-size heap buffer so that ASAN sees any overrun
    char *buf = (char*)malloc(tot);
    size_t o = 0;
    for (size_t j = 0; j < lens.size(); ++j) { off.push_back(o); len.push_back(lens[j]); sh.push_back(rand()%W); rid.push_back(j); for (int i=0;i<lens[j];++i) buf[o+i] = (rand()%50==0) ? 'N' : "ACGT"[rand()&3]; o += lens[j]; }
    int64_t cap = tot + 16; std::vector<uint64_t> out(2*cap); std::vector<int64_t> oo(lens.size()+1);
    emu_sketch_concurrency(it % 3 ? 0 : 100 + it);
    long n = emu_sketch_jobs(lens.size(), off.data(), len.data(), sh.data(), rid.data(), buf, w, k, Z, W, small, 3, oo.data(), out.data(), cap);
    printf("it %d k %d w %d Z %s small %d -> %ld\n", it, k, w, Z, small, n);
    if (small) {
      int maxlen = 0; for (int l : lens) if (l > maxlen) maxlen = l;
      int seg = maxlen / W * ones + ones + 1;
      for (int T = 32; T <= 128; T *= 2) { int pack = (T*8 - (w-1)) / seg; if (pack > 32) pack = 32; if (pack < 1) continue;
        int64_t stride = maxlen + 2; std::vector<uint64_t> o2(2*stride*lens.size()); std::vector<int32_t> cnt(lens.size());
        int rc = emu_sketch_packed(lens.size(), off.data(), len.data(), sh.data(), rid.data(), buf, w, k, Z, W, pack, T, 3, stride, cnt.data(), o2.data());
        printf("   packed T %d pack %d rc %d\n", T, pack, rc); }
    }
    free(buf);
  }
  return 0;
}
