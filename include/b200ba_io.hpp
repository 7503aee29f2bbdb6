// b200ba_io.hpp -- C++ readers / writers of the reference's on-disk formats either side of the
// bundle-adjustment path, over the containers of b200ba_shim.hpp: `dataset.bin` and the
// state directory (`intrinsicsN.yaml`, `rig_tr_global.yaml`, `camera_tr_rig.yaml`, `points.yaml`).
// The Python mirror is camera_calibration_b200/io.py; tests/test_cpp_io.py round-trips files between the two.
//
// Formats follow applications/camera_calibration/src/camera_calibration/io/calibration_io.cc (APP/io below):
//   * dataset.bin (SaveDataset :51-135, LoadDataset :137-246): magic "calib_data", u32 version 0, u32 camera
//     count, per camera u32 width, height; u32 imageset count, per imageset u32 filename length + bytes, per
//     camera u32 n + n x (f32 x, f32 y, i32 id); known geometries: u32 count, each f32 cell length, u32 n,
//     n x (i32 id, i32 x, i32 y). Integers are BIG-endian (htonl, io_util.h:56-64), floats raw host order
//     (io_util.h:66-69).
//   * camera model YAML (SaveCameraModel :526-647, LoadCameraModel :649-783): 14 significant digits, grids
//     flat row-major x, y, z; directions are re-normalised on load.
//   * poses YAML (SavePoses :785-839, LoadPoses :841-888): pose_count + list of index, tx ty tz, qx qy qz qw;
//     only used images are listed.
//   * points.yaml (:890-985): flat `points` + `feature_id_to_point_index` list.
// The reference parses YAML with yaml-cpp, which this repository must not depend on: the reader below
// understands the subset these files use (top-level `key : scalar`, `key : [flow, list]` possibly spanning
// lines, `key:` followed by a block list of flat maps, `#` comments).
//
// Header-only, host code only (no device work happens here). Every loader returns false on a malformed
// file, like the reference's.
#pragma once

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "b200ba_shim.hpp"

namespace b200ba_shim {
namespace io_detail {

inline uint32_t swap32(uint32_t v) {
  const uint16_t probe = 1;
  if (*reinterpret_cast<const uint8_t*>(&probe) == 0) return v;  // big-endian host
  return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
}
inline void put_u32(std::string* out, uint32_t v) {
  v = swap32(v);
  out->append(reinterpret_cast<const char*>(&v), 4);
}
inline void put_f32(std::string* out, float v) { out->append(reinterpret_cast<const char*>(&v), 4); }

struct Reader {
  const std::string& d;
  size_t pos = 0;
  bool ok = true;
  explicit Reader(const std::string& data) : d(data) {}
  bool need(size_t n) {
    if (!ok || d.size() - pos < n) ok = false;
    return ok;
  }
  uint32_t u32() {
    if (!need(4)) return 0;
    uint32_t v;
    std::memcpy(&v, d.data() + pos, 4);
    pos += 4;
    return swap32(v);
  }
  float f32() {
    if (!need(4)) return 0;
    float v;
    std::memcpy(&v, d.data() + pos, 4);
    pos += 4;
    return v;
  }
};

inline bool read_file(const std::string& path, std::string* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return true;
}
inline bool write_file(const std::string& path, const std::string& data) {
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f) return false;
  f.write(data.data(), static_cast<std::streamsize>(data.size()));
  return static_cast<bool>(f);
}
inline bool file_exists(const std::string& path) {
  struct stat st;
  return ::stat(path.c_str(), &st) == 0;
}
inline void make_directories(const std::string& path) {
  for (size_t i = 1; i <= path.size(); ++i)
    if (i == path.size() || path[i] == '/') {
      const std::string sub = path.substr(0, i);
      if (!sub.empty()) ::mkdir(sub.c_str(), 0777);
    }
}
inline std::string join(const std::string& dir, const std::string& name) {
  return (!dir.empty() && dir.back() == '/') ? dir + name : dir + "/" + name;
}

// std::ostream << double with setprecision(14)
inline std::string num(double v) {
  char buf[40];
  std::snprintf(buf, sizeof(buf), "%.14g", v);
  return buf;
}

// ---- the YAML subset -----------------------------------------------------------------------------------
struct Node {
  bool is_scalar = false, is_flow = false, is_block = false;
  std::string scalar;
  std::vector<double> flow;
  std::vector<std::map<std::string, std::string>> block;
};
using Document = std::map<std::string, Node>;

inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
  return s.substr(a, b - a);
}
inline bool to_double(const std::string& s, double* v) {
  const std::string t = trim(s);
  if (t.empty()) return false;
  if (t == ".nan" || t == ".NaN") { *v = std::nan(""); return true; }
  if (t == ".inf") { *v = HUGE_VAL; return true; }
  if (t == "-.inf") { *v = -HUGE_VAL; return true; }
  char* end = nullptr;
  errno = 0;
  *v = std::strtod(t.c_str(), &end);
  return end && *end == '\0';
}
inline bool to_int(const std::string& s, long long* v) {
  const std::string t = trim(s);
  if (t.empty()) return false;
  char* end = nullptr;
  errno = 0;
  *v = std::strtoll(t.c_str(), &end, 10);
  return end && *end == '\0' && errno == 0;
}
// splits "key : value" at the first ':' that is followed by a blank or ends the line
inline bool split_key(const std::string& line, std::string* key, std::string* value) {
  for (size_t i = 0; i < line.size(); ++i)
    if (line[i] == ':' && (i + 1 == line.size() || line[i + 1] == ' ' || line[i + 1] == '\t')) {
      *key = trim(line.substr(0, i));
      *value = trim(line.substr(i + 1));
      return !key->empty();
    }
  return false;
}
inline bool parse_flow(const std::string& text, std::vector<double>* out) {
  // text starts behind '[' and ends before ']'
  size_t a = 0;
  const std::string t = trim(text);
  if (t.empty()) return true;
  while (a <= t.size()) {
    size_t b = t.find(',', a);
    if (b == std::string::npos) b = t.size();
    double v;
    if (!to_double(t.substr(a, b - a), &v)) return false;
    out->push_back(v);
    a = b + 1;
  }
  return true;
}
inline bool parse_document(const std::string& text, Document* doc) {
  std::vector<std::string> lines;
  {
    std::istringstream ss(text);
    std::string l;
    while (std::getline(ss, l)) {
      // comments: a '#' at the start of the line or after a blank
      for (size_t i = 0; i < l.size(); ++i)
        if (l[i] == '#' && (i == 0 || l[i - 1] == ' ' || l[i - 1] == '\t')) {
          l.resize(i);
          break;
        }
      lines.push_back(l);
    }
  }
  size_t i = 0;
  while (i < lines.size()) {
    const std::string& raw = lines[i];
    if (trim(raw).empty()) { ++i; continue; }
    if (raw[0] == ' ' || raw[0] == '\t' || raw[0] == '-') return false;  // not a top-level key
    std::string key, value;
    if (!split_key(raw, &key, &value)) return false;
    Node node;
    ++i;
    if (!value.empty() && value[0] == '[') {
      std::string body = value.substr(1);
      while (body.find(']') == std::string::npos) {
        if (i >= lines.size()) return false;
        body += " " + lines[i++];
      }
      body.resize(body.find(']'));
      node.is_flow = true;
      if (!parse_flow(body, &node.flow)) return false;
    } else if (!value.empty()) {
      node.is_scalar = true;
      node.scalar = value;
    } else {
      // block list of flat maps (or nothing: an empty list)
      node.is_block = true;
      while (i < lines.size()) {
        const std::string t = trim(lines[i]);
        if (t.empty()) { ++i; continue; }
        if (lines[i][0] != ' ' && lines[i][0] != '\t' && lines[i][0] != '-') break;  // next top-level key
        std::string entry = t;
        if (entry[0] == '-') {
          node.block.emplace_back();
          entry = trim(entry.substr(1));
          if (entry.empty()) { ++i; continue; }
        }
        if (node.block.empty()) return false;
        std::string k, v;
        if (!split_key(entry, &k, &v)) return false;
        node.block.back()[k] = v;
        ++i;
      }
    }
    (*doc)[key] = node;
  }
  return true;
}
inline bool get_int(const Document& doc, const char* key, int* out) {
  auto it = doc.find(key);
  long long v;
  if (it == doc.end() || !it->second.is_scalar || !to_int(it->second.scalar, &v)) return false;
  *out = static_cast<int>(v);
  return true;
}
inline bool get_flow(const Document& doc, const char* key, const std::vector<double>** out) {
  auto it = doc.find(key);
  if (it == doc.end() || !it->second.is_flow) return false;
  *out = &it->second.flow;
  return true;
}
inline void append_flow(std::string* out, const double* v, size_t n) {
  out->push_back('[');
  for (size_t i = 0; i < n; ++i) {
    if (i) out->append(", ");
    out->append(num(v[i]));
  }
  out->append("]\n");
}
inline std::string dir_of(const std::string& path) {
  const size_t p = path.find_last_of('/');
  return p == std::string::npos ? std::string() : path.substr(0, p);
}

}  // namespace io_detail

// ---- dataset.bin ----------------------------------------------------------------------------------------
// APP/io:51-135
inline bool SaveDataset(const char* path, const Dataset& dataset) {
  using namespace io_detail;
  std::string out = "calib_data";
  put_u32(&out, 0);
  put_u32(&out, static_cast<uint32_t>(dataset.num_cameras()));
  for (int c = 0; c < dataset.num_cameras(); ++c) {
    put_u32(&out, static_cast<uint32_t>(dataset.GetImageSize(c).first));
    put_u32(&out, static_cast<uint32_t>(dataset.GetImageSize(c).second));
  }
  put_u32(&out, static_cast<uint32_t>(dataset.ImagesetCount()));
  for (int i = 0; i < dataset.ImagesetCount(); ++i) {
    std::shared_ptr<const Imageset> s = dataset.GetImageset(i);
    put_u32(&out, static_cast<uint32_t>(s->GetFilename().size()));
    out.append(s->GetFilename());
    for (int c = 0; c < dataset.num_cameras(); ++c) {
      const std::vector<PointFeature>& features = s->FeaturesOfCamera(c);
      put_u32(&out, static_cast<uint32_t>(features.size()));
      for (const PointFeature& f : features) {
        put_f32(&out, f.xy.x);
        put_f32(&out, f.xy.y);
        put_u32(&out, static_cast<uint32_t>(f.id));
      }
    }
  }
  put_u32(&out, static_cast<uint32_t>(dataset.known_geometries().size()));
  for (const KnownGeometry& g : dataset.known_geometries()) {
    put_f32(&out, g.cell_length_in_meters);
    put_u32(&out, static_cast<uint32_t>(g.feature_id_to_position.size()));
    for (const auto& item : g.feature_id_to_position) {
      put_u32(&out, static_cast<uint32_t>(item.first));
      put_u32(&out, static_cast<uint32_t>(item.second.first));
      put_u32(&out, static_cast<uint32_t>(item.second.second));
    }
  }
  make_directories(dir_of(path));
  return write_file(path, out);
}

// APP/io:137-246. `dataset` is replaced; false on a missing, truncated or foreign file.
inline bool LoadDataset(const char* path, std::shared_ptr<Dataset>* dataset) {
  using namespace io_detail;
  std::string data;
  if (!read_file(path, &data)) return false;
  if (data.size() < 10 || data.compare(0, 10, "calib_data") != 0) return false;
  Reader r(data);
  r.pos = 10;
  if (r.u32() != 0 || !r.ok) return false;  // version
  const uint32_t num_cameras = r.u32();
  if (!r.ok || num_cameras > (1u << 16)) return false;
  std::shared_ptr<Dataset> ds(new Dataset(static_cast<int>(num_cameras)));
  for (uint32_t c = 0; c < num_cameras; ++c) {
    const uint32_t w = r.u32(), h = r.u32();
    ds->SetImageSize(static_cast<int>(c), static_cast<int>(w), static_cast<int>(h));
  }
  const uint32_t num_imagesets = r.u32();
  for (uint32_t i = 0; r.ok && i < num_imagesets; ++i) {
    const uint32_t len = r.u32();
    if (!r.need(len)) return false;
    std::shared_ptr<Imageset> s = ds->NewImageset();
    s->SetFilename(data.substr(r.pos, len));
    r.pos += len;
    for (uint32_t c = 0; c < num_cameras; ++c) {
      const uint32_t n = r.u32();
      if (!r.need(static_cast<size_t>(n) * 12)) return false;
      std::vector<PointFeature>& features = s->FeaturesOfCamera(static_cast<int>(c));
      features.resize(n);
      for (uint32_t k = 0; k < n; ++k) {
        features[k].xy.x = r.f32();
        features[k].xy.y = r.f32();
        features[k].id = static_cast<int>(r.u32());
      }
    }
  }
  const uint32_t num_geometries = r.u32();
  for (uint32_t g = 0; r.ok && g < num_geometries; ++g) {
    KnownGeometry geometry;
    geometry.cell_length_in_meters = r.f32();
    const uint32_t n = r.u32();
    if (!r.need(static_cast<size_t>(n) * 12)) return false;
    for (uint32_t k = 0; k < n; ++k) {
      const int id = static_cast<int>(r.u32()), x = static_cast<int>(r.u32()), y = static_cast<int>(r.u32());
      geometry.feature_id_to_position.emplace_back(id, std::make_pair(x, y));
    }
    ds->known_geometries().push_back(geometry);
  }
  if (!r.ok) return false;
  *dataset = ds;
  return true;
}

// ---- camera models --------------------------------------------------------------------------------------
// APP/io:526-647
inline bool SaveCameraModel(CameraModel& model, const char* path) {
  using namespace io_detail;
  std::string out;
  auto header = [&](const char* type, bool area) {
    out += std::string("type : ") + type + "\n";
    out += "width : " + std::to_string(model.width()) + "\nheight : " + std::to_string(model.height()) + "\n";
    if (area) {
      out += "calibration_min_x : " + std::to_string(model.calibration_min_x()) + "\ncalibration_min_y : " +
             std::to_string(model.calibration_min_y()) + "\n";
      out += "calibration_max_x : " + std::to_string(model.calibration_max_x()) + "\ncalibration_max_y : " +
             std::to_string(model.calibration_max_y()) + "\n";
      int gw = 0, gh = 0;
      model.GetGridResolution(&gw, &gh);
      out += "grid_width : " + std::to_string(gw) + "\ngrid_height : " + std::to_string(gh) + "\n";
    }
  };
  const std::vector<double>& flat = model.flat_intrinsics();
  switch (model.type()) {
    case CameraModel::Type::CentralGeneric:
      header("CentralGenericModel", true);
      out += "# The grid is stored in row-major order, top to bottom. Each row is stored left to right. "
             "Each grid point is stored as x, y, z.\n";
      out += "grid : ";
      append_flow(&out, flat.data(), flat.size());
      break;
    case CameraModel::Type::NoncentralGeneric:
      header("NoncentralGenericModel", true);
      out += "# The grids are stored in row-major order, top to bottom. Each row is stored left to right. "
             "Each grid point is stored as x, y, z.\n";
      out += "point_grid : ";
      append_flow(&out, flat.data() + flat.size() / 2, flat.size() / 2);
      out += "direction_grid : ";
      append_flow(&out, flat.data(), flat.size() / 2);
      break;
    case CameraModel::Type::CentralOpenCV:
      header("CentralOpenCVModel", false);
      out += "parameters : ";
      append_flow(&out, flat.data(), flat.size());
      break;
    default:
      return false;  // model type not on the accelerated path
  }
  make_directories(dir_of(path));
  return write_file(path, out);
}

// APP/io:649-783. Returns an empty pointer on a malformed file or a model type that is not on this path.
inline std::shared_ptr<CameraModel> LoadCameraModel(const char* path) {
  using namespace io_detail;
  std::string text;
  Document doc;
  if (!read_file(path, &text) || !parse_document(text, &doc)) return nullptr;
  int width = 0, height = 0;
  if (!get_int(doc, "width", &width) || !get_int(doc, "height", &height) || width < 1 || height < 1) return nullptr;
  auto type_it = doc.find("type");
  if (type_it == doc.end() || !type_it->second.is_scalar) return nullptr;
  const std::string type = type_it->second.scalar;
  auto normalise = [](double* v, size_t n_points) {  // APP/io:672-675
    for (size_t i = 0; i < n_points; ++i) {
      const double norm = std::sqrt(v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
      v[3 * i] /= norm;
      v[3 * i + 1] /= norm;
      v[3 * i + 2] /= norm;
    }
  };
  if (type == "CentralGenericModel" || type == "NoncentralGenericModel") {
    int gw, gh, min_x, min_y, max_x, max_y;
    if (!get_int(doc, "grid_width", &gw) || !get_int(doc, "grid_height", &gh) || !get_int(doc, "calibration_min_x", &min_x) ||
        !get_int(doc, "calibration_min_y", &min_y) || !get_int(doc, "calibration_max_x", &max_x) ||
        !get_int(doc, "calibration_max_y", &max_y) || gw < 1 || gh < 1)
      return nullptr;
    const size_t n = 3 * static_cast<size_t>(gw) * gh;
    if (type == "CentralGenericModel") {
      const std::vector<double>* grid;
      if (!get_flow(doc, "grid", &grid) || grid->size() != n) return nullptr;
      std::shared_ptr<CentralGenericModel> m(new CentralGenericModel(gw, gh, min_x, min_y, max_x, max_y, width, height));
      m->grid = *grid;
      normalise(m->grid.data(), n / 3);
      return m;
    }
    const std::vector<double>*point_grid, *direction_grid;
    if (!get_flow(doc, "point_grid", &point_grid) || !get_flow(doc, "direction_grid", &direction_grid) ||
        point_grid->size() != n || direction_grid->size() != n)
      return nullptr;
    std::shared_ptr<NoncentralGenericModel> m(new NoncentralGenericModel(gw, gh, min_x, min_y, max_x, max_y, width, height));
    std::copy(direction_grid->begin(), direction_grid->end(), m->grids.begin());
    std::copy(point_grid->begin(), point_grid->end(), m->grids.begin() + n);
    normalise(m->grids.data(), n / 3);
    return m;
  }
  if (type == "CentralOpenCVModel") {
    const std::vector<double>* parameters;
    if (!get_flow(doc, "parameters", &parameters) || parameters->size() != 12) return nullptr;
    std::shared_ptr<CentralOpenCVModel> m(new CentralOpenCVModel(width, height));
    m->parameters = *parameters;
    return m;
  }
  return nullptr;
}

// ---- poses ----------------------------------------------------------------------------------------------
// APP/io:785-839 (also writes <path>.obj with the camera centres, like the reference)
inline bool SavePoses(const std::vector<bool>& image_used, const std::vector<SE3d>& poses, const char* path) {
  using namespace io_detail;
  if (image_used.size() != poses.size()) throw std::runtime_error("image_used and poses differ in size");  // CHECK_EQ :791
  std::string out =
      "# Each pose gives the B_tr_A transformation (i.e., A to B with right-multiplication), where the spaces A and B "
      "are defined by the filename. Quaternions are written as used by the Eigen library.\n";
  out += "pose_count: " + std::to_string(image_used.size()) + "\nposes:\n";
  std::string obj;
  for (size_t i = 0; i < poses.size(); ++i) {
    if (!image_used[i]) continue;
    const SE3d& T = poses[i];
    out += "  - index: " + std::to_string(i) + "\n    tx: " + num(T.tx) + "\n    ty: " + num(T.ty) + "\n    tz: " + num(T.tz) +
           "\n    qx: " + num(T.qx) + "\n    qy: " + num(T.qy) + "\n    qz: " + num(T.qz) + "\n    qw: " + num(T.qw) + "\n";
    // camera centre -R^T t
    SE3d inverse_rotation = T;
    inverse_rotation.qx = -T.qx; inverse_rotation.qy = -T.qy; inverse_rotation.qz = -T.qz;
    inverse_rotation.tx = inverse_rotation.ty = inverse_rotation.tz = 0;
    const Vec3d c = apply(inverse_rotation, Vec3d{-T.tx, -T.ty, -T.tz});
    obj += "v " + num(c.x) + " " + num(c.y) + " " + num(c.z) + " 1 0 0\n";
  }
  make_directories(dir_of(path));
  return write_file(path, out) && write_file(std::string(path) + ".obj", obj);
}

// APP/io:841-888
inline bool LoadPoses(std::vector<bool>* image_used, std::vector<SE3d>* poses, const char* path) {
  using namespace io_detail;
  std::string text;
  Document doc;
  if (!read_file(path, &text) || !parse_document(text, &doc)) return false;
  int count = 0;
  if (!get_int(doc, "pose_count", &count) || count < 0) return false;
  image_used->assign(count, false);
  poses->assign(count, SE3d());
  auto it = doc.find("poses");
  if (it == doc.end()) return true;
  if (!it->second.is_block) return false;
  for (const auto& item : it->second.block) {
    long long index;
    double v[7];
    const char* keys[7] = {"qw", "qx", "qy", "qz", "tx", "ty", "tz"};
    auto idx = item.find("index");
    if (idx == item.end() || !to_int(idx->second, &index) || index < 0 || index >= count) return false;
    for (int k = 0; k < 7; ++k) {
      auto f = item.find(keys[k]);
      if (f == item.end() || !to_double(f->second, &v[k])) return false;
    }
    const double norm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);  // SE3::setQuaternion normalises
    SE3d T;
    T.qw = v[0] / norm; T.qx = v[1] / norm; T.qy = v[2] / norm; T.qz = v[3] / norm;
    T.tx = v[4]; T.ty = v[5]; T.tz = v[6];
    (*image_used)[index] = true;
    (*poses)[index] = T;
  }
  return true;
}

// ---- points ---------------------------------------------------------------------------------------------
// APP/io:890-935. The mapping is written in ascending feature-id order (the reference iterates an
// unordered_map; the order carries no meaning).
inline bool SavePointsAndIndexMapping(const BAState& state, const char* path) {
  using namespace io_detail;
  std::string out = "# Each point is stored as x, y, z.\npoints : [";
  std::string obj;
  for (size_t i = 0; i < state.points.size(); ++i) {
    const Vec3d& p = state.points[i];
    if (i) out += ", ";
    out += num(p.x) + ", " + num(p.y) + ", " + num(p.z);
    obj += "v " + num(p.x) + " " + num(p.y) + " " + num(p.z) + " 0 0 1\n";
  }
  out += "]\nfeature_id_to_point_index:\n";
  std::map<int, int> ordered(state.feature_id_to_points_index.begin(), state.feature_id_to_points_index.end());
  for (const auto& item : ordered)
    out += "  - feature_id: " + std::to_string(item.first) + "\n    point_index: " + std::to_string(item.second) + "\n";
  make_directories(dir_of(path));
  return write_file(path, out) && write_file(std::string(path) + ".obj", obj);
}

// APP/io:937-985
inline bool LoadPointsAndIndexMapping(BAState* state, const char* path) {
  using namespace io_detail;
  std::string text;
  Document doc;
  if (!read_file(path, &text) || !parse_document(text, &doc)) return false;
  const std::vector<double>* flat;
  if (!get_flow(doc, "points", &flat) || flat->size() % 3 != 0) return false;
  std::vector<Vec3d> points(flat->size() / 3);
  for (size_t i = 0; i < points.size(); ++i) points[i] = Vec3d{(*flat)[3 * i], (*flat)[3 * i + 1], (*flat)[3 * i + 2]};
  std::unordered_map<int, int> mapping;
  auto it = doc.find("feature_id_to_point_index");
  if (it != doc.end()) {
    if (!it->second.is_block) return false;
    for (const auto& item : it->second.block) {
      long long id, index;
      auto a = item.find("feature_id"), b = item.find("point_index");
      if (a == item.end() || b == item.end() || !to_int(a->second, &id) || !to_int(b->second, &index)) return false;
      if (index < 0 || static_cast<size_t>(index) >= points.size()) return false;
      mapping[static_cast<int>(id)] = static_cast<int>(index);
    }
  }
  state->points.swap(points);
  state->feature_id_to_points_index.swap(mapping);
  return true;
}

// ---- the state directory --------------------------------------------------------------------------------
// APP/io:432-464
inline bool SaveBAState(const char* base_path, const BAState& state) {
  using namespace io_detail;
  make_directories(base_path);
  if (!SavePoses(state.image_used, state.rig_tr_global, join(base_path, "rig_tr_global.yaml").c_str())) return false;
  if (!SavePoses(std::vector<bool>(state.camera_tr_rig.size(), true), state.camera_tr_rig,
                 join(base_path, "camera_tr_rig.yaml").c_str()))
    return false;
  for (size_t c = 0; c < state.intrinsics.size(); ++c)
    if (!SaveCameraModel(*state.intrinsics[c], join(base_path, "intrinsics" + std::to_string(c) + ".yaml").c_str())) return false;
  return SavePointsAndIndexMapping(state, join(base_path, "points.yaml").c_str());
}

// APP/io:466-523. With a dataset, the features' point indices are refreshed from the loaded mapping.
inline bool LoadBAState(const char* base_path, BAState* state, Dataset* dataset) {
  using namespace io_detail;
  BAState loaded;
  if (!LoadPoses(&loaded.image_used, &loaded.rig_tr_global, join(base_path, "rig_tr_global.yaml").c_str())) return false;
  std::vector<bool> all_used;
  if (!LoadPoses(&all_used, &loaded.camera_tr_rig, join(base_path, "camera_tr_rig.yaml").c_str())) return false;
  for (int c = 0;; ++c) {
    const std::string path = join(base_path, "intrinsics" + std::to_string(c) + ".yaml");
    if (!file_exists(path)) {
      if (c == 0) return false;
      break;
    }
    std::shared_ptr<CameraModel> model = LoadCameraModel(path.c_str());
    if (!model) return false;
    loaded.intrinsics.push_back(model);
  }
  if (!LoadPointsAndIndexMapping(&loaded, join(base_path, "points.yaml").c_str())) return false;
  if (dataset) {
    // unknown feature ids make the reference's .at() throw: report a malformed pair instead
    for (int i = 0; i < dataset->ImagesetCount(); ++i)
      for (int c = 0; c < dataset->num_cameras(); ++c)
        for (const PointFeature& f : dataset->GetImageset(i)->FeaturesOfCamera(c))
          if (!loaded.feature_id_to_points_index.count(f.id)) return false;
    loaded.ComputeFeatureIdToPointsIndex(dataset);
  }
  *state = loaded;
  return true;
}

}  // namespace b200ba_shim
