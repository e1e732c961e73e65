// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Stand-in for Cello's umbrella header, written for this repository so that
// the reference's *unmodified* VL+CT hot-path sources (read in place from
// /root/reference/src, never copied) compile into oracle/_ref/libvlct_ref.so
// without Charm++, HDF5, the parameter-file parser or the rest of Cello.
//
// It supplies light fakes for the small surface of Cello that those sources
// touch: Block / Data / Field / FieldDescr / Grouping (field lookup by name over
// caller-provided memory), Method / Physics / Compute / Refresh (empty PUP::able
// bases), ParameterGroup (a string->string dictionary) and the cello:: accessors.
// The real CelloView / ViewMap / error-macro headers are used as they are.
#ifndef VLCT_SHIM_CELLO_HPP
#define VLCT_SHIM_CELLO_HPP

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <cstdlib>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include <charm++.h>
#include "pup_stl.h"

// --- real reference headers (macros + views), found via -I .../src/Cello ----
#include "cello_defines.hpp"   // FORCE_INLINE
#include "error_Error.hpp"     // ASSERT*/ERROR*/WARNING* macros
#include "view_CelloView.hpp"
#include "view_ViewCollec.hpp"
#include "view_StringIndRdOnlyMap.hpp"
#include "view_ViewMap.hpp"

#ifndef MIN
#define MIN(a,b) (((a)<(b))?(a):(b))
#endif
#ifndef MAX
#define MAX(a,b) (((a)>(b))?(a):(b))
#endif
#define UNUSED(x) (void)(x)

enum precision_enum {
  precision_unknown, precision_default, precision_single, precision_double,
  precision_extended80, precision_extended96, precision_quadruple
};
typedef int precision_type;

enum struct ghost_choice { exclude, include, permit };
enum class InitCycleKind { fresh, charmrestart, fresh_or_noncharm_restart };

enum { ir_post_unset = -1 };

// src/Cello/cello.hpp:94-95 (flat index of a C-ordered (z, y, x) array)
inline int INDEX(int ix, int iy, int iz, int nx, int ny)
{ return ix + nx * (iy + ny * iz); }

// src/Cello/cello.hpp:111-124
enum face_enum { face_lower = 0, face_upper = 1, face_all };
enum axis_enum { axis_x = 0, axis_y = 1, axis_z = 2, axis_all };

//----------------------------------------------------------------------

class Grouping {
public:
  void add(const std::string& item, const std::string& group)
  { groups_[group].push_back(item); }
  int size(const std::string& group) const {
    auto it = groups_.find(group);
    return (it == groups_.end()) ? 0 : (int) it->second.size();
  }
  std::string item(const std::string& group, int index) const
  { return groups_.at(group).at(index); }
  std::vector<std::string> group_list(const std::string& group) const {
    auto it = groups_.find(group);
    return (it == groups_.end()) ? std::vector<std::string>() : it->second;
  }
  bool is_in(const std::string& item, const std::string& group) const {
    auto it = groups_.find(group);
    if (it == groups_.end()) return false;
    return std::find(it->second.begin(), it->second.end(), item)
      != it->second.end();
  }
private:
  std::map<std::string, std::vector<std::string>> groups_;
};

//----------------------------------------------------------------------

/// One entry per permanent field: a name, a centering and caller-owned memory
struct ShimFieldEntry {
  std::string name;
  int cx, cy, cz;      // 1 if face-centred along that axis
  double* ptr;         // caller-owned storage, (mz+cz, my+cy, mx+cx) C-order
};

class FieldDescr {
public:
  FieldDescr() : gx_(0), gy_(0), gz_(0) {}
  int insert(const std::string& name, int cx, int cy, int cz) {
    entries_.push_back({name, cx, cy, cz, nullptr});
    return (int) entries_.size() - 1;
  }
  void clear() { entries_.clear(); groups_ = Grouping(); }
  void set_ghost_depth(int gx, int gy, int gz) { gx_=gx; gy_=gy; gz_=gz; }

  bool is_field(const std::string& name) const { return field_id(name) >= 0; }
  int field_id(const std::string& name) const {
    for (std::size_t i = 0; i < entries_.size(); i++)
      if (entries_[i].name == name) return (int) i;
    return -1;
  }
  std::string field_name(int id) const { return entries_.at(id).name; }
  int field_count() const { return (int) entries_.size(); }
  void ghost_depth(int /*id*/, int* gx, int* gy, int* gz) const
  { if (gx) *gx = gx_; if (gy) *gy = gy_; if (gz) *gz = gz_; }
  void centering(int id, int* cx, int* cy, int* cz) const {
    const ShimFieldEntry& e = entries_.at(id);
    if (cx) *cx = e.cx; if (cy) *cy = e.cy; if (cz) *cz = e.cz;
  }
  int precision(int /*id*/) const { return precision_default; }
  Grouping* groups() { return &groups_; }
  const Grouping* groups() const { return &groups_; }
  const std::vector<ShimFieldEntry>& entries() const { return entries_; }
private:
  std::vector<ShimFieldEntry> entries_;
  Grouping groups_;
  int gx_, gy_, gz_;
};

//----------------------------------------------------------------------

class FieldData {
public:
  FieldData() : nx(0), ny(0), nz(0) {}
  int nx, ny, nz;                 // active-zone size
  std::vector<double*> ptrs;      // one per FieldDescr entry
};

class Field {
public:
  Field() : descr_(nullptr), data_(nullptr) {}
  Field(FieldDescr* d, FieldData* f) : descr_(d), data_(f) {}

  int field_id(const std::string& name) const { return descr_->field_id(name); }
  bool is_field(const std::string& name) const { return descr_->is_field(name);}
  std::string field_name(int id) const { return descr_->field_name(id); }
  void ghost_depth(int id, int* gx, int* gy, int* gz) const
  { descr_->ghost_depth(id, gx, gy, gz); }
  void centering(int id, int* cx, int* cy, int* cz) const
  { descr_->centering(id, cx, cy, cz); }
  int precision(int id) const { return descr_->precision(id); }
  int field_count() const { return descr_->field_count(); }
  /// src/Cello/data_FieldData.cpp:254-264
  void cell_width(double xm, double xp, double* hx, double ym = 0, double yp = 0,
                  double* hy = 0, double zm = 0, double zp = 0, double* hz = 0) const {
    if (hx) *hx = (xp - xm) / data_->nx;
    if (hy) *hy = (yp - ym) / data_->ny;
    if (hz) *hz = (zp - zm) / data_->nz;
  }
  bool is_temporary(int) const { return false; }   // only permanent fields here
  char* values(int id) { return reinterpret_cast<char*>(data_->ptrs.at(id)); }
  void size(int* nx, int* ny, int* nz) const
  { if (nx) *nx = data_->nx; if (ny) *ny = data_->ny; if (nz) *nz = data_->nz; }
  void dimensions(int id, int* mx, int* my, int* mz) const {
    int gx, gy, gz, cx, cy, cz;
    descr_->ghost_depth(id, &gx, &gy, &gz);
    descr_->centering(id, &cx, &cy, &cz);
    if (mx) *mx = data_->nx + 2*gx + cx;
    if (my) *my = data_->ny + 2*gy + cy;
    if (mz) *mz = data_->nz + 2*gz + cz;
  }
  Grouping* groups() { return descr_->groups(); }
  bool ghosts_allocated() const { return true; }
  double history_time(int /*history*/) const { return 0.0; }
  const char* values(int id, int /*history*/ = 0) const
  { return (id < 0) ? nullptr : (const char*) data_->ptrs.at(id); }
  const char* values(const std::string& name, int history = 0) const
  { return values(field_id(name), history); }

  template<class T>
  CelloView<T,3> view(int id, ghost_choice choice = ghost_choice::include,
                      int history = 0) {
    static_assert(std::is_same<T,double>::value, "oracle shim is fp64 only");
    if (id < 0 || choice == ghost_choice::exclude || history != 0) {
      ERROR("Field::view (shim)", "unsupported view request");
    }
    int mx, my, mz;
    dimensions(id, &mx, &my, &mz);
    return CelloView<T,3>(data_->ptrs.at(id), mz, my, mx);
  }
  template<class T>
  CelloView<T,3> view(const std::string& name,
                      ghost_choice choice = ghost_choice::include,
                      int history = 0) {
    int id = field_id(name);
    if (id < 0) { ERROR1("Field::view (shim)", "no field named %s", name.c_str()); }
    return view<T>(id, choice, history);
  }
  template<class T>
  CelloView<const T,3> view(int id, ghost_choice choice = ghost_choice::include,
                            int history = 0) const
  { return const_cast<Field*>(this)->view<T>(id, choice, history); }
  template<class T>
  CelloView<const T,3> view(const std::string& name,
                            ghost_choice choice = ghost_choice::include,
                            int history = 0) const
  { return const_cast<Field*>(this)->view<T>(name, choice, history); }
private:
  FieldDescr* descr_;
  FieldData* data_;
};

//----------------------------------------------------------------------

/// Minimal stand-ins for Cello's flux-correction containers
/// (src/Cello/data_FaceFluxes.hpp, data_FluxData.hpp): one array per
/// (field, axis, side); indexing ix + mx*(iy + my*iz), with extent 1 along the
/// face's normal (data_FaceFluxes.hpp:98-122).
class FaceFluxes {
public:
  FaceFluxes(int mx = 1, int my = 1, int mz = 1)
    : mx_(mx), my_(my), mz_(mz), fluxes_((std::size_t) mx * my * mz, 0.0) {}
  int get_size(int* pmx = nullptr, int* pmy = nullptr, int* pmz = nullptr) const {
    if (pmx) *pmx = mx_;
    if (pmy) *pmy = my_;
    if (pmz) *pmz = mz_;
    return mx_ * my_ * mz_;
  }
  double* flux_array(int* dx = nullptr, int* dy = nullptr, int* dz = nullptr) {
    if (dx) *dx = 1;
    if (dy) *dy = mx_;
    if (dz) *dz = mx_ * my_;
    return fluxes_.data();
  }
private:
  int mx_, my_, mz_;
  std::vector<double> fluxes_;
};
class FluxData {
public:
  int num_fields() const { return (int) field_list_.size(); }
  int index_field(int i_f) const { return field_list_.at(i_f); }
  FaceFluxes* block_fluxes(int axis, int face, int i_f)
  { return &fluxes_.at(((std::size_t) i_f * 3 + axis) * 2 + face); }
  void allocate(int nx, int ny, int nz, std::vector<int> field_list, bool) {
    field_list_ = field_list;
    fluxes_.clear();
    for (std::size_t i_f = 0; i_f < field_list.size(); i_f++)
      for (int axis = 0; axis < 3; axis++)
        for (int face = 0; face < 2; face++)
          fluxes_.push_back(FaceFluxes(axis == 0 ? 1 : nx, axis == 1 ? 1 : ny,
                                       axis == 2 ? 1 : nz));
  }
private:
  std::vector<int> field_list_;
  std::vector<FaceFluxes> fluxes_;
};

class Data {
public:
  Data(FieldDescr* d) : descr_(d) {}
  Field field() { return Field(descr_, &field_data); }
  FluxData* flux_data() { return &flux_data_; }
  void lower(double* x, double* y, double* z) const
  { if (x) *x = xm[0]; if (y) *y = xm[1]; if (z) *z = xm[2]; }
  void field_cell_width(double* hx, double* hy, double* hz) const
  { if (hx) *hx = h[0]; if (hy) *hy = h[1]; if (hz) *hz = h[2]; }
  /// positions of cell centres (c = 0) or faces (c = 1) along each axis, ghost
  /// zones included: lower + (index + 0.5 (1 - c)) * width, the expression of
  /// Data::field_cell_faces (src/Cello/data_Data.cpp:91-121)
  void upper(double* x, double* y, double* z) const {
    const int n[3] = { field_data.nx, field_data.ny, field_data.nz };
    double* out[3] = { x, y, z };
    for (int a = 0; a < 3; a++) if (out[a]) *out[a] = xm[a] + n[a] * h[a];
  }
  void field_cells(double* x, double* y, double* z, int gx = 0, int gy = 0,
                   int gz = 0) const
  { field_cell_faces(x, y, z, gx, gy, gz, 0, 0, 0); }
  void field_cell_faces(double* x, double* y, double* z, int gx = 0, int gy = 0,
                        int gz = 0, int cx = 1, int cy = 1, int cz = 1) const {
    const int n[3] = { field_data.nx, field_data.ny, field_data.nz };
    const int g[3] = { gx, gy, gz }, c[3] = { cx, cy, cz };
    double* out[3] = { x, y, z };
    for (int a = 0; a < 3; a++) {
      const double d = (c[a] == 0) ? 0.5 : 0;
      for (int i = -g[a]; i < n[a] + g[a] + c[a]; i++)
        out[a][i + g[a]] = xm[a] + (i + d) * h[a];
    }
  }
  FieldData field_data;
  double xm[3] = {0,0,0};
  double h[3] = {1,1,1};
private:
  FieldDescr* descr_;
  FluxData flux_data_;
};

class Index {
public:
  bool is_root() const { return true; }
};

/// In the shim every block is an EnzoBlock (see Enzo/enzo.hpp)
class Block {
public:
  Block(FieldDescr* d) : data_(d), dt_(0), time_(0), cycle_(0),
                         compute_done_count(0) {}
  virtual ~Block() {}
  Data* data() { return &data_; }
  bool is_leaf() const { return true; }
  double dt() const { return dt_; }
  double time() const { return time_; }
  int cycle() const { return cycle_; }
  Index index() const { return Index(); }
  void cell_width(double* hx, double* hy, double* hz) const
  { const_cast<Data&>(data_).field_cell_width(hx, hy, hz); }
  void compute_done() { compute_done_count++; }
  void initial_done() {}
  void set_dt(double dt) { dt_ = dt; }
  void set_cycle(int cycle) { cycle_ = cycle; }
  Data data_;
  double dt_, time_;
  int cycle_;
  int compute_done_count;
};

//----------------------------------------------------------------------

class Refresh {
public:
  void add_all_fields() {}
  void add_field(const std::string&) {}
  void add_field(int) {}
};

/// src/Cello/mesh_Hierarchy.hpp:178-179: blocks on this process
class Hierarchy {
public:
  Hierarchy() : num_blocks_(1) {}
  size_t num_blocks() const throw() { return num_blocks_; }
  void set_num_blocks(size_t n) { num_blocks_ = n; }
private:
  size_t num_blocks_;
};

class Simulation {
public:
  void refresh_set_name(int, const std::string&) {}
  Hierarchy* hierarchy() const throw() { return const_cast<Hierarchy*>(&hierarchy_); }
private:
  Hierarchy hierarchy_;
};

class Monitor {
public:
  void print(const char*, const char*, ...) {}
};

class Problem {
public:
  bool method_exists(const std::string&) const { return false; }
  bool method_precedes(const std::string&, const std::string&) const
  { return false; }
};

/// A dictionary standing in for the parameter-file group "Method:mhd_vlct"
class ParameterGroup {
public:
  ParameterGroup() : path_("Method:mhd_vlct") {}
  void set(const std::string& key, const std::string& value)
  { values_[key] = value; }
  const std::string* param(const std::string& key) const {
    auto it = values_.find(key);
    return (it == values_.end()) ? nullptr : &it->second;
  }
  std::string value_string(const std::string& key,
                           const std::string& deflt) const
  { const std::string* p = param(key); return p ? *p : deflt; }
  double value_float(const std::string& key, double deflt) const
  { const std::string* p = param(key); return p ? atof(p->c_str()) : deflt; }
  bool value_logical(const std::string& key, bool deflt) const {
    const std::string* p = param(key);
    return p ? (*p == "true" || *p == "1") : deflt;
  }
  int value_integer(const std::string& key, int deflt) const
  { const std::string* p = param(key); return p ? atoi(p->c_str()) : deflt; }
  /// list parameters are stored as "v0,v1,v2"
  int list_length(const std::string& key) const {
    const std::string* p = param(key);
    if (p == nullptr || p->empty()) return 0;
    return 1 + (int) std::count(p->begin(), p->end(), ',');
  }
  double list_value_float(int index, const std::string& key,
                          double deflt = 0.0) const {
    const std::string* p = param(key);
    if (p == nullptr) return deflt;
    std::size_t pos = 0;
    for (int i = 0; i < index; i++) {
      pos = p->find(',', pos);
      if (pos == std::string::npos) return deflt;
      pos++;
    }
    return atof(p->c_str() + pos);
  }
  std::string full_name(const std::string& key) const
  { return path_ + ":" + key; }
  std::string get_group_path() const { return path_; }
private:
  std::string path_;
  std::map<std::string, std::string> values_;
};

//----------------------------------------------------------------------

namespace cello {
  const double pi = 3.14159265358979324;

  int rank();
  FieldDescr* field_descr();
  Simulation* simulation();
  Refresh* refresh(int ir);
  Monitor* monitor();
  bool is_initial_cycle(InitCycleKind kind) noexcept;
}

//----------------------------------------------------------------------

class Method : public PUP::able {
public:
  Method(double courant = 1.0) : ir_post_(0), courant_(courant) {}
  Method(CkMigrateMessage* m) : PUP::able(m), ir_post_(0), courant_(1.0) {}
  virtual ~Method() {}
  virtual void pup(PUP::er& p) { PUP::able::pup(p); }
  virtual void compute(Block* block) throw() = 0;
  virtual std::string name() throw() = 0;
  virtual double timestep(Block*) throw()
  { return std::numeric_limits<double>::max(); }
  double courant() const throw() { return courant_; }
  void set_courant(double courant) throw() { courant_ = courant; }
protected:
  int ir_post_;
  double courant_;
};

/// src/Cello/problem_Mask.hpp: never instantiated here (boundaries without masks)
class Mask {
public:
  virtual ~Mask() {}
  virtual bool evaluate(double t, double x, double y, double z) const = 0;
};

/// src/Cello/problem_Boundary.hpp:14-120: what a Boundary subclass needs
class Boundary : public PUP::able {
public:
  Boundary() throw() : axis_(axis_all), face_(face_all), mask_(nullptr) {}
  Boundary(axis_enum axis, face_enum face, std::shared_ptr<Mask> mask) throw()
    : axis_(axis), face_(face), mask_(mask) {}
  Boundary(CkMigrateMessage* m)
    : PUP::able(m), axis_(axis_all), face_(face_all), mask_(nullptr) {}
  virtual ~Boundary() throw() {}
  virtual void pup(PUP::er& p) { PUP::able::pup(p); }
  virtual void enforce(Block* block, face_enum face = face_all,
                       axis_enum axis = axis_all) const throw() = 0;
protected:
  bool applies_(axis_enum axis, face_enum face) const throw()
  { return ((axis_ == axis_all || axis == axis_) &&
            (face_ == face_all || face == face_)); }
  axis_enum axis_;
  face_enum face_;
  std::shared_ptr<Mask> mask_;
};

/// Inert stand-ins for the parameter-file machinery (flex / bison generated in
/// the reference): no parameter exists, so EnzoInitialBCenter holds no Value
/// expressions and only its static array helpers do any work here.
enum parameter_type_shim { parameter_unknown, parameter_list };
class Parameters {
public:
  void group_set(int, const std::string&) {}
  int type(const std::string&) const { return parameter_unknown; }
  int list_length(const std::string&) const { return 0; }
  void pup(PUP::er&) {}
};
class Value {
public:
  Value(Parameters*, const std::string&) {}
  template <class T>
  void evaluate(T*, double, int, int, double*, int, int, double*, int, int,
                double*) const {}
};

/// src/Cello/problem_Initial.hpp: only what an Initial subclass needs to compile
class Initial : public PUP::able {
public:
  Initial(int cycle, double time) throw() : cycle_(cycle), time_(time) {}
  Initial(CkMigrateMessage* m) : PUP::able(m), cycle_(0), time_(0.0) {}
  virtual ~Initial() {}
  virtual void pup(PUP::er& p) { PUP::able::pup(p); }
  virtual void enforce_block(Block* block, const Hierarchy* hierarchy) throw() = 0;
protected:
  int cycle_;
  double time_;
};

class Physics : public PUP::able {
public:
  Physics() {}
  Physics(CkMigrateMessage* m) : PUP::able(m) {}
  virtual ~Physics() {}
  virtual std::string type() const = 0;
};

class Compute : public PUP::able {
public:
  Compute() : i_hist_(0) {}
  Compute(CkMigrateMessage* m) : PUP::able(m), i_hist_(0) {}
  virtual ~Compute() {}
  virtual void compute(Block* block) throw() = 0;
  virtual int  get_history(int) { return i_hist_; }
  virtual void set_history(int i_hist) { i_hist_ = i_hist; }
protected:
  int i_hist_;
};

#endif /* VLCT_SHIM_CELLO_HPP */
