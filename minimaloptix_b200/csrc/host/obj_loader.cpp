#include "obj_loader.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace tinyobj {
namespace {

inline bool isDigit(char c) { return (unsigned)(c - '0') < 10u; }
inline bool isBlank(char c) { return c == ' ' || c == '\t'; }
inline bool isEol(char c) { return c == '\r' || c == '\n' || c == '\0'; }

struct VIdx { int v, vt, vn; };

// 1-based -> 0-based; negative counts back from the current array size; 0 is illegal.
bool fixIndex(int idx, int n, int* out) {
  if (idx > 0) { *out = idx - 1; return true; }
  if (idx < 0) { *out = n + idx; return true; }
  return false;
}

// "i", "i/j", "i//k", "i/j/k"
bool parseTriple(const char*& p, int nv, int nvn, int nvt, VIdx* out) {
  VIdx vi{-1, -1, -1};
  if (!fixIndex(atoi(p), nv, &vi.v)) return false;
  p += strcspn(p, "/ \t\r");
  if (*p != '/') { *out = vi; return true; }
  ++p;
  if (*p == '/') {
    ++p;
    if (!fixIndex(atoi(p), nvn, &vi.vn)) return false;
    p += strcspn(p, "/ \t\r");
    *out = vi;
    return true;
  }
  if (!fixIndex(atoi(p), nvt, &vi.vt)) return false;
  p += strcspn(p, "/ \t\r");
  if (*p != '/') { *out = vi; return true; }
  ++p;
  if (!fixIndex(atoi(p), nvn, &vi.vn)) return false;
  p += strcspn(p, "/ \t\r");
  *out = vi;
  return true;
}

real_t parseReal(const char*& p, double dflt = 0.0) {
  p += strspn(p, " \t");
  const char* end = p + strcspn(p, " \t\r");
  double val = dflt;
  tryParseDouble(p, end, &val);
  p = end;
  return (real_t)val;
}

// Even-odd point-in-polygon (W. R. Franklin's pnpoly), as tinyobj uses for the ear test.
bool pointInTri(const real_t* vx, const real_t* vy, real_t tx, real_t ty) {
  bool c = false;
  for (int i = 0, j = 2; i < 3; j = i++) {
    if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) c = !c;
  }
  return c;
}

void emitTri(mesh_t& m, const VIdx& a, const VIdx& b, const VIdx& c) {
  m.indices.push_back({a.v, a.vn, a.vt});
  m.indices.push_back({b.v, b.vn, b.vt});
  m.indices.push_back({c.v, c.vn, c.vt});
  m.num_face_vertices.push_back(3);
  m.material_ids.push_back(-1);
}

// Append one polygon to the mesh, ear-clipping it when triangulate is set.
void exportFace(mesh_t& m, const std::vector<VIdx>& face, bool triangulate, const std::vector<real_t>& v) {
  size_t np = face.size();
  if (np < 3) return;
  if (!triangulate) {
    for (const VIdx& k : face) m.indices.push_back({k.v, k.vn, k.vt});
    m.num_face_vertices.push_back((unsigned char)np);
    m.material_ids.push_back(-1);
    return;
  }
  // Projection axes: drop the axis along which the first non-degenerate corner's normal is largest.
  size_t ax[2] = {1, 2};
  for (size_t k = 0; k < np; ++k) {
    size_t a = (size_t)face[k % np].v, b = (size_t)face[(k + 1) % np].v, c = (size_t)face[(k + 2) % np].v;
    if (3 * a + 2 >= v.size() || 3 * b + 2 >= v.size() || 3 * c + 2 >= v.size()) continue;
    real_t e0x = v[3 * b] - v[3 * a], e0y = v[3 * b + 1] - v[3 * a + 1], e0z = v[3 * b + 2] - v[3 * a + 2];
    real_t e1x = v[3 * c] - v[3 * b], e1y = v[3 * c + 1] - v[3 * b + 1], e1z = v[3 * c + 2] - v[3 * b + 2];
    real_t cx = std::fabs(e0y * e1z - e0z * e1y);
    real_t cy = std::fabs(e0z * e1x - e0x * e1z);
    real_t cz = std::fabs(e0x * e1y - e0y * e1x);
    const real_t eps = std::numeric_limits<real_t>::epsilon();
    if (cx > eps || cy > eps || cz > eps) {
      if (!(cx > cy && cx > cz)) {
        ax[0] = 0;
        if (cz > cx && cz > cy) ax[1] = 1;
      }
      break;
    }
  }
  real_t area = 0;
  for (size_t k = 0; k < np; ++k) {
    size_t a = (size_t)face[k].v, b = (size_t)face[(k + 1) % np].v;
    if (3 * a + ax[0] >= v.size() || 3 * a + ax[1] >= v.size() || 3 * b + ax[0] >= v.size() || 3 * b + ax[1] >= v.size())
      continue;
    area += (v[3 * a + ax[0]] * v[3 * b + ax[1]] - v[3 * a + ax[1]] * v[3 * b + ax[0]]) * (real_t)0.5;
  }
  std::vector<VIdx> rest = face;
  size_t guess = 0, budget = face.size(), prevCount = rest.size();
  while (rest.size() > 3 && budget > 0) {
    np = rest.size();
    if (guess >= np) guess -= np;
    if (prevCount != np) { prevCount = np; budget = np; } else { budget--; }
    VIdx ind[3];
    real_t vx[3], vy[3];
    for (int k = 0; k < 3; ++k) {
      ind[k] = rest[(guess + k) % np];
      size_t vi = (size_t)ind[k].v;
      if (3 * vi + ax[0] >= v.size() || 3 * vi + ax[1] >= v.size()) { vx[k] = 0; vy[k] = 0; }
      else { vx[k] = v[3 * vi + ax[0]]; vy[k] = v[3 * vi + ax[1]]; }
    }
    real_t cr = (vx[1] - vx[0]) * (vy[2] - vy[1]) - (vy[1] - vy[0]) * (vx[2] - vx[1]);
    if (cr * area < (real_t)0.0) { guess++; continue; }  // reflex corner
    bool overlap = false;
    for (size_t o = 3; o < np; ++o) {
      size_t ovi = (size_t)rest[(guess + o) % np].v;
      if (3 * ovi + ax[0] >= v.size() || 3 * ovi + ax[1] >= v.size()) continue;
      if (pointInTri(vx, vy, v[3 * ovi + ax[0]], v[3 * ovi + ax[1]])) { overlap = true; break; }
    }
    if (overlap) { guess++; continue; }
    emitTri(m, ind[0], ind[1], ind[2]);
    rest.erase(rest.begin() + (long)((guess + 1) % np));
  }
  if (rest.size() == 3) emitTri(m, rest[0], rest[1], rest[2]);
}

bool flushGroup(shape_t& shape, const std::vector<std::vector<VIdx>>& faces, bool hadLines, const std::string& name,
                bool triangulate, const std::vector<real_t>& v) {
  if (faces.empty() && !hadLines) return false;
  if (!faces.empty()) {
    for (const auto& f : faces) exportFace(shape.mesh, f, triangulate, v);
    shape.name = name;
  }
  return true;
}

}  // namespace

// Grammar: [sign] digits ['.' digits] [('e'|'E') [sign] digits].  The value is accumulated
// digit by digit in double (integer part: m = 10 m + d; fraction: m += d * 10^-k with a small
// table for k < 8 and pow() beyond), then scaled by 5^e via pow and 2^e via ldexp.  This is
// tinyobj's algorithm; it is deliberately NOT correctly rounded, and we need its exact bits.
bool tryParseDouble(const char* s, const char* s_end, double* result) {
  if (s >= s_end) return false;
  double mant = 0.0;
  int expo = 0;
  char sign = '+', esign = '+';
  const char* c = s;
  int nread = 0;
  if (*c == '+' || *c == '-') { sign = *c; ++c; }
  else if (!isDigit(*c)) return false;
  bool more = c != s_end;
  while (more && isDigit(*c)) {
    mant *= 10;
    mant += (int)(*c - '0');
    ++c; ++nread;
    more = c != s_end;
  }
  if (nread == 0) return false;
  auto finish = [&]() {
    *result = (sign == '+' ? 1 : -1) * (expo ? std::ldexp(mant * std::pow(5.0, expo), expo) : mant);
    return true;
  };
  if (!more) return finish();
  if (*c == '.') {
    ++c;
    nread = 1;
    more = c != s_end;
    static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
    while (more && isDigit(*c)) {
      mant += (int)(*c - '0') * (nread < 8 ? lut[nread] : std::pow(10.0, -nread));
      ++nread; ++c;
      more = c != s_end;
    }
  } else if (*c != 'e' && *c != 'E') {
    return finish();
  }
  if (!more) return finish();
  if (*c == 'e' || *c == 'E') {
    ++c;
    more = c != s_end;
    if (more && (*c == '+' || *c == '-')) { esign = *c; ++c; }
    else if (!isDigit(*c)) return false;
    nread = 0;
    more = c != s_end;
    while (more && isDigit(*c)) {
      if (expo < 100000) expo = expo * 10 + (int)(*c - '0');   // saturate: tinyobj's int overflows (UB) on "1e9999999999"; the value is inf or 0 either way
      ++c; ++nread;
      more = c != s_end;
    }
    if (esign == '-') expo = -expo;
    if (nread == 0) return false;
  }
  return finish();
}

bool LoadObj(attrib_t* attrib, std::vector<shape_t>* shapes, std::vector<material_t>* materials, std::string* warn,
             std::string* err, const char* filename, const char* /*mtl_basedir*/, bool triangulate) {
  attrib->vertices.clear(); attrib->normals.clear(); attrib->texcoords.clear();
  shapes->clear();
  if (materials) materials->clear();
  FILE* fp = fopen(filename, "rb");
  if (!fp) {
    if (err) *err += std::string("Cannot open file [") + filename + "]\n";
    return false;
  }
  std::string data;
  {
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) data.append(buf, n);
    fclose(fp);
  }
  std::vector<real_t> v, vn, vt;
  std::vector<std::vector<VIdx>> faces;
  bool hadLines = false;
  std::string name;
  shape_t shape;
  size_t lineNo = 0, pos = 0;
  std::string line;
  while (pos < data.size()) {
    size_t e = data.find('\n', pos);
    if (e == std::string::npos) e = data.size();
    line.assign(data, pos, e - pos);
    pos = e + 1;
    ++lineNo;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    const char* p = line.c_str();
    p += strspn(p, " \t");
    if (*p == '\0' || *p == '#') continue;
    if (p[0] == 'v' && isBlank(p[1])) {
      p += 2;
      real_t x = parseReal(p), y = parseReal(p), z = parseReal(p);
      v.push_back(x); v.push_back(y); v.push_back(z);
      continue;
    }
    if (p[0] == 'v' && p[1] == 'n' && isBlank(p[2])) {
      p += 3;
      real_t x = parseReal(p), y = parseReal(p), z = parseReal(p);
      vn.push_back(x); vn.push_back(y); vn.push_back(z);
      continue;
    }
    if (p[0] == 'v' && p[1] == 't' && isBlank(p[2])) {
      p += 3;
      real_t x = parseReal(p), y = parseReal(p);
      vt.push_back(x); vt.push_back(y);
      continue;
    }
    if (p[0] == 'l' && isBlank(p[1])) { hadLines = true; continue; }
    if (p[0] == 'f' && isBlank(p[1])) {
      p += 2;
      p += strspn(p, " \t");
      std::vector<VIdx> face;
      face.reserve(4);
      while (!isEol(*p)) {
        VIdx vi;
        if (!parseTriple(p, (int)(v.size() / 3), (int)(vn.size() / 3), (int)(vt.size() / 2), &vi)) {
          if (err) *err += "Failed parse `f' line(e.g. zero value for face index. line " + std::to_string(lineNo) + ".)\n";
          return false;
        }
        face.push_back(vi);
        p += strspn(p, " \t\r");
      }
      faces.push_back(std::move(face));
      continue;
    }
    if (p[0] == 'g' && isBlank(p[1])) {
      flushGroup(shape, faces, hadLines, name, triangulate, v);
      if (!shape.mesh.indices.empty()) shapes->push_back(shape);
      shape = shape_t();
      faces.clear();
      hadLines = false;
      // tokens after 'g', joined by single spaces
      std::vector<std::string> names;
      while (!isEol(*p)) {
        p += strspn(p, " \t");
        size_t n = strcspn(p, " \t\r");
        names.emplace_back(p, n);
        p += n;
        p += strspn(p, " \t\r");
      }
      if (names.size() < 2) {
        if (warn) { *warn += "Empty group name. line: " + std::to_string(lineNo) + "\n"; name.clear(); }
      } else {
        name = names[1];
        for (size_t i = 2; i < names.size(); ++i) name += " " + names[i];
      }
      continue;
    }
    if (p[0] == 'o' && isBlank(p[1])) {
      if (flushGroup(shape, faces, hadLines, name, triangulate, v)) shapes->push_back(shape);
      faces.clear();
      hadLines = false;
      shape = shape_t();
      name = p + 2;
      continue;
    }
    // usemtl / mtllib / s / t and anything else: no effect on geometry.
  }
  bool ret = flushGroup(shape, faces, hadLines, name, triangulate, v);
  if (ret || !shape.mesh.indices.empty()) shapes->push_back(shape);
  attrib->vertices.swap(v);
  attrib->normals.swap(vn);
  attrib->texcoords.swap(vt);
  return true;
}

}  // namespace tinyobj
