// ./ppp_main workpiece.pcd [out_prefix]   — the reference's ./main flow (src/main.cpp:6-36) on the
// B200 path: construct path_generater(file, 15), estimate_normal(), Contact_Path_Generation()
// (show() needs a display and is not part of the hot path).  With a second argument the normals
// and ordered contour nodes are dumped as raw binaries for the parity tests.
// ./ppp_main --sect config.txt workpiece.pcd [out_prefix] runs the SectPath / config.txt flow
// (src/connect.cpp:23-26 with Dynamic_adjustment off).
#include <cstdio>
#include <cstring>
#include <string>

#include "ppp_adapter.h"

static void dump(const std::string& path, const void* p, size_t bytes) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return;
  if (bytes) fwrite(p, 1, bytes, f);
  fclose(f);
}

static void dump_contours(const std::string& prefix, const ppp_host::Contours& c) {
  dump(prefix + ".off.i64", c.offsets.data(), c.offsets.size() * 8);
  dump(prefix + ".y.f64", c.y.data(), c.y.size() * 8);
  dump(prefix + ".x.f64", c.x.data(), c.x.size() * 8);
  dump(prefix + ".z.f64", c.z.data(), c.z.size() * 8);
}

int main(int argc, char** argv) {
  try {
    if (argc >= 4 && !strcmp(argv[1], "--sect")) {
      SectPath sp(argv[2], argv[3]);
      sp.estimate_normal();
      sp.GenPath();
      if (argc >= 5) {
        dump(std::string(argv[4]) + ".normals.f32", sp.normal_cloud->points.data(), sp.normal_cloud->points.size() * sizeof(pcl::Normal));
        dump_contours(argv[4], sp.last_contours());
      }
      return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "--slicing")) {  // the sweep main.cpp:28 leaves commented out
      path_generater pg(argv[2], 15);
      pg.slicing_method();
      if (argc >= 4) dump_contours(argv[3], pg.last_contours());
      return 0;
    }
    if (argc < 2) {
      printf("Usage: %s workpiece.pcd [out_prefix] | --sect config.txt workpiece.pcd [out_prefix]\n", argv[0]);
      return -1;
    }
    double Radius = 15;  // src/main.cpp:23
    path_generater pg(argv[1], Radius);
    pg.estimate_normal();
    pg.Contact_Path_Generation();
    if (argc >= 3) {
      dump(std::string(argv[2]) + ".normals.f32", pg.cloud_with_normals->points.data(), pg.cloud_with_normals->points.size() * sizeof(pcl::Normal));
      dump_contours(argv[2], pg.last_contours());
    }
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "ppp_main: %s\n", e.what());
    return 1;
  }
}
