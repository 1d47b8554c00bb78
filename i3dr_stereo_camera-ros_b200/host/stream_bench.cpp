// Drives B200FrameStream from one process over N GPUs (BASELINE config c4: a stream of 2448x2048x256d pairs dealt over the
// GPUs of one box, no NCCL).
//   stream_bench <pairs.raw> <distinct> <W> <H> <frames> <gpus> <lanes> minD D window uniq speckleRange speckleSize cap p1 p2 mode
// pairs.raw holds `distinct` pairs (left image, right image, ...) of W x H bytes; the stream cycles through them.  One untimed
// pass, then `frames` frames timed with the host clock around submit..drain (the copies from / to page-locked host buffers are
// inside).  Prints one JSON line: frames/s, the CRC-32 of every distinct frame's disparity (last occurrence in the stream) and
// the GPU that produced it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "stream_driver.h"

static uint32_t crc32(const void *data, size_t n)
{
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  const unsigned char *p = static_cast<const unsigned char *>(data);
  for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

int main(int argc, char **argv)
{
  if (argc < 18) {
    std::cerr << "usage: stream_bench pairs.raw distinct W H frames gpus lanes minD D window uniq speckleRange speckleSize cap p1 p2 mode" << std::endl;
    return 2;
  }
  const int distinct = atoi(argv[2]), W = atoi(argv[3]), H = atoi(argv[4]), frames = atoi(argv[5]), gpus = atoi(argv[6]), lanes = atoi(argv[7]);
  b200sgm_params p;
  p.minDisparity = atoi(argv[8]); p.numDisparities = atoi(argv[9]); p.blockSize = atoi(argv[10]); p.uniquenessRatio = atoi(argv[11]);
  p.speckleRange = atoi(argv[12]); p.speckleWindowSize = atoi(argv[13]); p.preFilterCap = atoi(argv[14]); p.P1 = atoi(argv[15]);
  p.P2 = atoi(argv[16]); p.mode = atoi(argv[17]); p.disp12MaxDiff = 0;
  if (distinct < 1 || W < 1 || H < 1 || frames < 1 || gpus < 1) return 2;
  const size_t npix = size_t(W) * H;
  uint8_t *images = nullptr;
  int16_t *out = nullptr;
  // output buffers cycle over a window of stream positions that is a multiple of the GPU count (a buffer always returns to the
  // same GPU) and longer than what can be pending there (3 * lanes, see B200FrameStream::submit)
  int window = gpus * (3 * lanes + 1);
  while (window < distinct) window += gpus;
  if (b200sgm_host_alloc(2 * npix * distinct, reinterpret_cast<void **>(&images)) != 0 ||
      b200sgm_host_alloc(npix * 2 * window, reinterpret_cast<void **>(&out)) != 0) {
    std::cerr << "page-locked allocation failed" << std::endl;
    return 1;
  }
  {
    std::ifstream f(argv[1], std::ios::binary);
    f.read(reinterpret_cast<char *>(images), 2 * npix * distinct);
    if (!f) { std::cerr << "cannot read " << argv[1] << std::endl; return 2; }
  }
  std::vector<int> devs;
  for (int g = 0; g < gpus; g++) devs.push_back(g);
  B200FrameStream stream(devs, W, H, p, lanes);
  if (!stream.ok()) { std::cerr << stream.error() << std::endl; return 1; }
  auto pass = [&](int n) {
    for (int i = 0; i < n; i++) {
      const int k = i % distinct;
      stream.submit(uint64_t(i), images + 2 * npix * k, images + 2 * npix * k + npix, out + npix * (i % window));
    }
    return stream.drain();
  };
  int rc = pass(std::min(frames, 2 * gpus * lanes));   // warm-up
  if (rc < 0) { std::cerr << stream.error() << std::endl; return 1; }
  const auto t0 = std::chrono::steady_clock::now();
  rc = pass(frames);
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc < 0) { std::cerr << stream.error() << std::endl; return 1; }
  printf("{\"frames\": %d, \"gpus\": %d, \"lanes\": %d, \"seconds\": %.6f, \"frames_per_s\": %.3f, \"status\": %d, \"crc\": [", frames, gpus, lanes, s,
         frames / s, rc);
  for (int k = 0; k < distinct; k++) {
    // last stream position of distinct frame k
    int i = frames - 1;
    while (i >= 0 && i % distinct != k) i--;
    printf("%s\"%08x\"", k ? ", " : "", i >= 0 ? crc32(out + npix * (i % window), npix * 2) : 0u);
  }
  printf("], \"owner_of_last\": %d}\n", stream.owner(uint64_t(frames - 1)));
  b200sgm_host_free(images);
  b200sgm_host_free(out);
  return 0;
}
