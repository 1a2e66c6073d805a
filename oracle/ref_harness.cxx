// ref_harness.cxx -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference (Panda3D 1.11.0 built from /root/reference by
// oracle/build_ref.sh into oracle/_ref/) through its public C++ API:
//   * `run  <recipe.p3cs> <golden.p3cs>`: builds real GeomVertexData /
//     TransformBlendTable / SliderTable objects from a recipe, calls
//     GeomVertexData::animate_vertices(true, thread) per frame
//     (panda/src/gobj/geomVertexData.cxx:912) and writes the flattened inputs
//     (as the reference objects report them) plus the output array bytes.
//   * `egg  <model.egg> <anim.egg> <golden.p3cs> <frame>...`: the same for the
//     first animated Geom of an Actor-style model (config C1).
//   * `bench <recipe.p3cs> <iters> <warmup>`: times animate_vertices with all
//     transforms / sliders dirtied per iteration; prints one JSON line.
// It is the pin of oracle/skin_oracle.c and the CPU baseline ("kind":"reference").
#include "pandabase.h"
#include "geomVertexData.h"
#include "geomVertexFormat.h"
#include "geomVertexArrayFormat.h"
#include "geomVertexArrayData.h"
#include "transformBlendTable.h"
#include "transformBlend.h"
#include "transformTable.h"
#include "geomVertexReader.h"
#include "geomVertexAnimationSpec.h"
#include "userVertexTransform.h"
#include "userVertexSlider.h"
#include "sliderTable.h"
#include "internalName.h"
#include "sparseArray.h"
#include "thread.h"
#include "load_egg_file.h"
#include "pandaNode.h"
#include "nodePath.h"
#include "nodePathCollection.h"
#include "geomNode.h"
#include "geom.h"
#include "character.h"
#include "auto_bind.h"
#include "animControlCollection.h"
#include "coordinateSystem.h"
#include "load_prc_file.h"
#include "characterJoint.h"
#include "characterJointBundle.h"
#include "jointVertexTransform.h"
#include "animChannelMatrixXfmTable.h"
#include "partBundle.h"
#include "animBundleNode.h"
#include "animChannelScalarTable.h"
#include "characterSlider.h"
#include "characterVertexSlider.h"
#include "animControl.h"
#include "movingPartMatrix.h"
#include "transformState.h"
#include "animChannel.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

// ---------------------------------------------------------------- blob file
enum { DT_U8 = 0, DT_I32 = 1, DT_U32 = 2, DT_F32 = 3, DT_U16 = 4, DT_I64 = 5, DT_F64 = 6 };
struct Blob {
  uint32_t dtype = DT_U8;
  std::vector<uint64_t> shape;
  std::vector<unsigned char> data;
  const int32_t *i32() const { return (const int32_t *)data.data(); }
  const float *f32() const { return (const float *)data.data(); }
  size_t count() const { size_t n = 1; for (auto s : shape) n *= s; return n; }
};
typedef std::map<std::string, Blob> Case;

static bool read_case(const char *path, Case &c) {
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); return false; }
  char magic[8]; uint32_t ver, cnt;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "P3CSBLOB", 8) != 0) { fclose(f); return false; }
  if (fread(&ver, 4, 1, f) != 1 || fread(&cnt, 4, 1, f) != 1) { fclose(f); return false; }
  for (uint32_t i = 0; i < cnt; ++i) {
    uint32_t nl; if (fread(&nl, 4, 1, f) != 1) break;
    std::string name(nl, '\0');
    if (nl && fread(&name[0], 1, nl, f) != nl) break;
    long pad = (4 - (nl & 3)) & 3; fseek(f, pad, SEEK_CUR);
    Blob b; uint32_t nd; uint64_t nb;
    if (fread(&b.dtype, 4, 1, f) != 1 || fread(&nd, 4, 1, f) != 1) break;
    b.shape.resize(nd);
    if (nd && fread(b.shape.data(), 8, nd, f) != nd) break;
    if (fread(&nb, 8, 1, f) != 1) break;
    b.data.resize(nb);
    if (nb && fread(b.data.data(), 1, nb, f) != nb) break;
    fseek(f, (8 - (nb & 7)) & 7, SEEK_CUR);
    c[name] = std::move(b);
  }
  fclose(f);
  return true;
}

static bool write_case(const char *path, const Case &c) {
  FILE *f = fopen(path, "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", path); return false; }
  uint32_t ver = 1, cnt = (uint32_t)c.size();
  fwrite("P3CSBLOB", 1, 8, f); fwrite(&ver, 4, 1, f); fwrite(&cnt, 4, 1, f);
  static const char zeros[8] = {0};
  for (auto &kv : c) {
    uint32_t nl = (uint32_t)kv.first.size();
    fwrite(&nl, 4, 1, f); fwrite(kv.first.data(), 1, nl, f); fwrite(zeros, 1, (4 - (nl & 3)) & 3, f);
    const Blob &b = kv.second; uint32_t nd = (uint32_t)b.shape.size(); uint64_t nb = b.data.size();
    fwrite(&b.dtype, 4, 1, f); fwrite(&nd, 4, 1, f);
    if (nd) fwrite(b.shape.data(), 8, nd, f);
    fwrite(&nb, 8, 1, f);
    if (nb) fwrite(b.data.data(), 1, nb, f);
    fwrite(zeros, 1, (8 - (nb & 7)) & 7, f);
  }
  fclose(f);
  return true;
}

template <class T> static Blob make_blob(uint32_t dt, const std::vector<T> &v, std::vector<uint64_t> shape) {
  Blob b; b.dtype = dt; b.shape = shape; b.data.resize(v.size() * sizeof(T));
  if (!v.empty()) memcpy(b.data.data(), v.data(), b.data.size());
  return b;
}
static Blob str_blob(const std::string &s) {
  Blob b; b.dtype = DT_U8; b.shape = {s.size()}; b.data.assign(s.begin(), s.end()); return b;
}
static std::vector<std::string> split_lines(const Blob &b) {
  std::vector<std::string> out; std::string cur;
  for (unsigned char ch : b.data) { if (ch == '\n') { out.push_back(cur); cur.clear(); } else cur.push_back((char)ch); }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

// ------------------------------------------------- export of reference objects
static std::vector<int32_t> sparse_ranges(const SparseArray &s) {
  std::vector<int32_t> r;
  nassertr(!s.is_inverse(), r);
  for (size_t i = 0; i < s.get_num_subranges(); ++i) { r.push_back(s.get_subrange_begin(i)); r.push_back(s.get_subrange_end(i)); }
  return r;
}

struct Exported {
  std::vector<const VertexTransform *> palette;  // ascending address = palette order
  std::vector<int> out_arrays;                   // source array index of each output array
  CPT(SliderTable) slider_table;                 // the mesh's SliderTable (may be null)
};

// Flattens a source GeomVertexData into the `case` sections, reading every fact
// back from the reference objects (formats, tables), not from any recipe.
static Exported export_vertex_data(const GeomVertexData *vd, Case &c) {
  Exported ex;
  ex.slider_table = vd->get_slider_table();
  const GeomVertexFormat *fmt = vd->get_format();
  CPT(GeomVertexFormat) post = fmt->get_post_animated_format();
  int num_rows = vd->get_num_rows();
  int na = fmt->get_num_arrays();

  std::vector<int32_t> columns; std::string names;
  std::vector<int32_t> strides, is_output;
  int col_count = 0;
  for (int a = 0; a < na; ++a) {
    const GeomVertexArrayFormat *af = fmt->get_array(a);
    strides.push_back(af->get_stride());
    int kept = 0;
    for (int k = 0; k < af->get_num_columns(); ++k) {
      const GeomVertexColumn *col = af->get_column(k);
      columns.insert(columns.end(), {a, col->get_start(), col->get_num_values(), (int)col->get_numeric_type(), (int)col->get_contents()});
      names += col->get_name()->get_name(); names += '\n';
      ++col_count;
      if (col->get_name() != InternalName::get_transform_blend() && col->get_contents() != GeomEnums::C_morph_delta) ++kept;
    }
    is_output.push_back(kept > 0 ? 1 : 0);
    if (kept > 0) ex.out_arrays.push_back(a);
    CPT(GeomVertexArrayDataHandle) h = vd->get_array(a)->get_handle();
    const unsigned char *p = h->get_read_pointer(true);
    Blob b; b.dtype = DT_U8; b.shape = {(uint64_t)num_rows, (uint64_t)af->get_stride()};
    b.data.assign(p, p + (size_t)num_rows * af->get_stride());
    c["src_array" + std::to_string(a)] = std::move(b);
  }
  nassertr((int)ex.out_arrays.size() == post->get_num_arrays(), ex);
  c["columns"] = make_blob(DT_I32, columns, {(uint64_t)col_count, 5});
  c["column_names"] = str_blob(names);
  c["src_strides"] = make_blob(DT_I32, strides, {(uint64_t)na});
  c["src_is_output"] = make_blob(DT_I32, is_output, {(uint64_t)na});

  // skin columns: points then vectors of the post-animated format (geomVertexData.cxx:1582,1622)
  std::vector<int32_t> skin; int nskin = 0;
  for (size_t pass = 0; pass < 2; ++pass) {
    size_t n = pass == 0 ? post->get_num_points() : post->get_num_vectors();
    for (size_t i = 0; i < n; ++i) {
      const InternalName *name = pass == 0 ? post->get_point(i) : post->get_vector(i);
      int a = fmt->get_array_with(name);
      const GeomVertexColumn *col = fmt->get_column(name);
      skin.insert(skin.end(), {a, col->get_start(), col->get_num_values(), (int)col->get_numeric_type(), (int)col->get_contents()});
      ++nskin;
    }
  }
  c["skin_columns"] = make_blob(DT_I32, skin, {(uint64_t)nskin, 5});

  // blend table
  CPT(TransformBlendTable) tbt = vd->get_transform_blend_table();
  std::vector<int32_t> bcol = {-1, 0, 0, 0, 0};
  std::vector<int32_t> boff = {0}, bxf; std::vector<float> bw; std::vector<int32_t> rows;
  if (tbt != nullptr) {
    int ba = fmt->get_array_with(InternalName::get_transform_blend());
    if (ba >= 0) {
      const GeomVertexColumn *col = fmt->get_column(InternalName::get_transform_blend());
      bcol = {ba, col->get_start(), col->get_num_values(), (int)col->get_numeric_type(), (int)col->get_contents()};
    }
    for (size_t b = 0; b < tbt->get_num_blends(); ++b) {
      const TransformBlend &bl = tbt->get_blend(b);
      for (size_t k = 0; k < bl.get_num_transforms(); ++k) ex.palette.push_back(bl.get_transform(k));
    }
    std::sort(ex.palette.begin(), ex.palette.end());
    ex.palette.erase(std::unique(ex.palette.begin(), ex.palette.end()), ex.palette.end());
    for (size_t b = 0; b < tbt->get_num_blends(); ++b) {
      const TransformBlend &bl = tbt->get_blend(b);
      for (size_t k = 0; k < bl.get_num_transforms(); ++k) {  // stored order = accumulation order
        int idx = (int)(std::lower_bound(ex.palette.begin(), ex.palette.end(), bl.get_transform(k)) - ex.palette.begin());
        bxf.push_back(idx); bw.push_back(bl.get_weight(k));
      }
      boff.push_back((int32_t)bxf.size());
    }
    rows = sparse_ranges(tbt->get_rows());
  }
  c["blend_column"] = make_blob(DT_I32, bcol, {5});
  c["blend_offsets"] = make_blob(DT_I32, boff, {(uint64_t)boff.size()});
  c["blend_transform"] = make_blob(DT_I32, bxf, {(uint64_t)bxf.size()});
  c["blend_weight"] = make_blob(DT_F32, bw, {(uint64_t)bw.size()});
  c["rows"] = make_blob(DT_I32, rows, {(uint64_t)rows.size() / 2, 2});

  // sliders + morph applications in the reference's own iteration order
  // (geomVertexData.cxx:1467-1480)
  CPT(SliderTable) st = vd->get_slider_table();
  std::vector<int32_t> morphs, mroff = {0}, mrows; int nm = 0; int ns = 0;
  if (st != nullptr) {
    ns = (int)st->get_num_sliders();
    for (int mi = 0; mi < (int)fmt->get_num_morphs(); ++mi) {
      const InternalName *sname = fmt->get_morph_slider(mi);
      const InternalName *bname = fmt->get_morph_base(mi);
      const InternalName *dname = fmt->get_morph_delta(mi);
      const SparseArray &sl = st->find_sliders(sname);
      if (sl.is_zero()) continue;
      const GeomVertexColumn *bc = fmt->get_column(bname), *dc = fmt->get_column(dname);
      for (size_t sni = 0; sni < sl.get_num_subranges(); ++sni) {
        for (int sn = sl.get_subrange_begin(sni); sn < sl.get_subrange_end(sni); ++sn) {
          morphs.insert(morphs.end(), {sn,
            fmt->get_array_with(bname), bc->get_start(), bc->get_num_values(), (int)bc->get_numeric_type(), (int)bc->get_contents(),
            fmt->get_array_with(dname), dc->get_start(), dc->get_num_values(), (int)dc->get_numeric_type(), (int)dc->get_contents()});
          std::vector<int32_t> r = sparse_ranges(st->get_slider_rows(sn));
          mrows.insert(mrows.end(), r.begin(), r.end());
          mroff.push_back((int32_t)mrows.size() / 2);
          ++nm;
        }
      }
    }
  }
  c["morphs"] = make_blob(DT_I32, morphs, {(uint64_t)nm, 11});
  c["morph_row_offsets"] = make_blob(DT_I32, mroff, {(uint64_t)mroff.size()});
  c["morph_rows"] = make_blob(DT_I32, mrows, {(uint64_t)mrows.size() / 2, 2});

  std::vector<int32_t> meta = {num_rows, na, (int)ex.out_arrays.size(), (int)ex.palette.size(),
                               tbt != nullptr ? (int)tbt->get_num_blends() : 0, ns,
                               (int)get_default_coordinate_system(), 0};
  c["meta"] = make_blob(DT_I32, meta, {(uint64_t)meta.size()});
  return ex;
}

static void append_bytes(Blob &b, const unsigned char *p, size_t n) { b.data.insert(b.data.end(), p, p + n); }

// Records one animated frame: palette, slider values and output array bytes.
static void record_frame(const GeomVertexData *vd, const Exported &ex, Case &c, int frame_index) {
  Thread *th = Thread::get_current_thread();
  Blob &pal = c["palette"]; pal.dtype = DT_F32;
  for (const VertexTransform *t : ex.palette) {
    LMatrix4 m; t->get_matrix(m);
    LMatrix4f mf = LCAST(float, m);
    append_bytes(pal, (const unsigned char *)mf.get_data(), 64);
  }
  pal.shape = {(uint64_t)frame_index + 1, (uint64_t)ex.palette.size(), 16};
  CPT(SliderTable) st = vd->get_slider_table();
  Blob &sv = c["sliders"]; sv.dtype = DT_F32;
  size_t ns = st != nullptr ? st->get_num_sliders() : 0;
  for (size_t s = 0; s < ns; ++s) { float v = (float)st->get_slider(s)->get_slider(); append_bytes(sv, (const unsigned char *)&v, 4); }
  sv.shape = {(uint64_t)frame_index + 1, (uint64_t)ns};

  CPT(GeomVertexData) out = vd->animate_vertices(true, th);
  nassertv(out != vd);
  nassertv(out->get_num_arrays() == (int)ex.out_arrays.size());
  for (size_t o = 0; o < ex.out_arrays.size(); ++o) {
    Blob &e = c["expect" + std::to_string(o)]; e.dtype = DT_U8;
    CPT(GeomVertexArrayDataHandle) h = out->get_array(o)->get_handle();
    size_t nbytes = (size_t)out->get_num_rows() * out->get_format()->get_array(o)->get_stride();
    append_bytes(e, h->get_read_pointer(true), nbytes);
    e.shape = {(uint64_t)frame_index + 1, nbytes};
  }
  // the blended matrices, as TransformBlend::get_blend reports them
  CPT(TransformBlendTable) tbt = vd->get_transform_blend_table();
  if (tbt != nullptr) {
    Blob &bm = c["expect_blends"]; bm.dtype = DT_F32;
    for (size_t b = 0; b < tbt->get_num_blends(); ++b) {
      LMatrix4 m; tbt->get_blend(b).update_blend(th); tbt->get_blend(b).get_blend(m, th);
      LMatrix4f mf = LCAST(float, m);
      append_bytes(bm, (const unsigned char *)mf.get_data(), 64);
    }
    bm.shape = {(uint64_t)frame_index + 1, (uint64_t)tbt->get_num_blends(), 16};
  }
  ((int32_t *)c["meta"].data.data())[7] = frame_index + 1;
}

// ------------------------------------------------- build objects from a recipe
struct Built {
  PT(GeomVertexData) vd;
  std::vector<PT(UserVertexTransform)> transforms;  // palette order == ascending address
  std::vector<PT(UserVertexSlider)> sliders;
};

static bool build_from_recipe(const Case &c, Built &out) {
  const int32_t *meta = c.at("meta").i32();
  int num_rows = meta[0], na = meta[1], T = meta[3], B = meta[4], S = meta[5];
  std::vector<std::string> names = split_lines(c.at("column_names"));
  const Blob &cols = c.at("columns");
  const int32_t *strides = c.at("src_strides").i32();
  PT(GeomVertexFormat) f = new GeomVertexFormat;
  for (int a = 0; a < na; ++a) {
    PT(GeomVertexArrayFormat) af = new GeomVertexArrayFormat;
    for (size_t k = 0; k < cols.shape[0]; ++k) {
      const int32_t *cc = cols.i32() + k * 5;
      if (cc[0] != a) continue;
      // the recipe lists get_num_values(); the packed types hold 4 (3: packed_ufloat) values per component
      const int per_component = (cc[3] == GeomEnums::NT_packed_dcba || cc[3] == GeomEnums::NT_packed_dabc) ? 4
                                : cc[3] == GeomEnums::NT_packed_ufloat ? 3 : 1;
      af->add_column(InternalName::make(names[k]), cc[2] / per_component, (GeomEnums::NumericType)cc[3], (GeomEnums::Contents)cc[4], cc[1], 1);
    }
    if (af->get_stride() < strides[a]) af->set_stride(strides[a]);
    if (af->get_stride() != strides[a]) { fprintf(stderr, "array %d: stride %d != recipe %d\n", a, af->get_stride(), strides[a]); return false; }
    f->add_array(af);
  }
  GeomVertexAnimationSpec spec; spec.set_panda(); f->set_animation(spec);
  CPT(GeomVertexFormat) fmt = GeomVertexFormat::register_format(f);
  out.vd = new GeomVertexData("recipe", fmt, GeomEnums::UH_static);
  out.vd->unclean_set_num_rows(num_rows);
  for (int a = 0; a < na; ++a) {
    const Blob &b = c.at("src_array" + std::to_string(a));
    if ((int)fmt->get_array(a)->get_stride() != strides[a] || b.data.size() != (size_t)num_rows * strides[a]) { fprintf(stderr, "array %d size mismatch\n", a); return false; }
    PT(GeomVertexArrayDataHandle) h = out.vd->modify_array_handle(a);
    memcpy(h->get_write_pointer(), b.data.data(), b.data.size());
  }
  // transforms, palette index == ascending pointer order (the order TransformBlend stores entries in,
  // transformBlend.I:354-357)
  out.transforms.resize(T);
  for (int j = 0; j < T; ++j) out.transforms[j] = new UserVertexTransform("j" + std::to_string(j));
  std::sort(out.transforms.begin(), out.transforms.end(), [](const PT(UserVertexTransform) &a, const PT(UserVertexTransform) &b) { return a.p() < b.p(); });
  if (c.at("blend_column").i32()[0] >= 0) {
    PT(TransformBlendTable) tbt = new TransformBlendTable;
    const int32_t *off = c.at("blend_offsets").i32(), *xf = c.at("blend_transform").i32();
    const float *w = c.at("blend_weight").f32();
    for (int b = 0; b < B; ++b) {
      TransformBlend bl;
      for (int k = off[b]; k < off[b + 1]; ++k) bl.add_transform(out.transforms[xf[k]], w[k]);
      size_t idx = tbt->add_blend(bl);
      if ((int)idx != b) { fprintf(stderr, "blend %d deduplicated to %d: recipe must hold unique blends\n", b, (int)idx); return false; }
    }
    const Blob &rows = c.at("rows");
    SparseArray sa;
    for (size_t i = 0; i < rows.shape[0]; ++i) sa.set_range(rows.i32()[2 * i], rows.i32()[2 * i + 1] - rows.i32()[2 * i]);
    tbt->set_rows(sa);
    out.vd->set_transform_blend_table(tbt);
  }
  if (S > 0) {
    std::vector<std::string> snames = split_lines(c.at("slider_names"));
    const int32_t *sroff = c.at("slider_row_offsets").i32(), *srows = c.at("slider_rows").i32();
    PT(SliderTable) st = new SliderTable;
    for (int s = 0; s < S; ++s) {
      PT(UserVertexSlider) sl = new UserVertexSlider(snames[s]);
      SparseArray sa;
      for (int i = sroff[s]; i < sroff[s + 1]; ++i) sa.set_range(srows[2 * i], srows[2 * i + 1] - srows[2 * i]);
      st->add_slider(sl, sa);
      out.sliders.push_back(sl);
    }
    out.vd->set_slider_table(SliderTable::register_table(st));
  }
  return true;
}

static void set_frame(Built &b, const Case &c, int frame) {
  const int32_t *meta = c.at("meta").i32();
  int T = meta[3], S = meta[5];
  const float *pal = c.at("palette").f32() + (size_t)frame * T * 16;
  for (int j = 0; j < T; ++j) {
    const float *m = pal + j * 16;
    LMatrix4f mf(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
    b.transforms[j]->set_matrix(LCAST(PN_stdfloat, mf));
  }
  if (S > 0) {
    const float *sv = c.at("sliders").f32() + (size_t)frame * S;
    for (int s = 0; s < S; ++s) b.sliders[s]->set_slider(sv[s]);
  }
}

// The reference takes its hpr convention from the PRC variable `coordinate-system`
// (panda/src/linmath/coordinateSystem.cxx:30); a recipe may ask for a non-default one.
static void apply_recipe_coordinate_system(const Case &in) {
  static const char *names[] = {"default", "zup-right", "yup-right", "zup-left", "yup-left"};
  int cs = in.at("meta").i32()[6];
  if (cs >= 1 && cs <= 4) load_prc_file_data("recipe", std::string("coordinate-system ") + names[cs]);
}

static int cmd_run(const char *in_path, const char *out_path) {
  Case in; if (!read_case(in_path, in)) return 1;
  apply_recipe_coordinate_system(in);
  Built b; if (!build_from_recipe(in, b)) return 1;
  int frames = (int)in.at("palette").shape[0];
  Case out;
  Exported ex = export_vertex_data(b.vd, out);
  // palette order of the export must equal the recipe's (both ascending address)
  for (size_t j = 0; j < ex.palette.size(); ++j) {
    if (ex.palette[j] != b.transforms[j].p() && ex.palette.size() == b.transforms.size()) { fprintf(stderr, "palette order mismatch\n"); return 1; }
  }
  if (ex.palette.size() != b.transforms.size()) {
    // some recipe transforms are unused: export the full recipe palette instead so indices agree
    fprintf(stderr, "recipe has unused transforms (%d used of %d): not supported\n", (int)ex.palette.size(), (int)b.transforms.size());
    return 1;
  }
  for (int fr = 0; fr < frames; ++fr) { set_frame(b, in, fr); record_frame(b.vd, ex, out, fr); }
  for (const char *k : {"slider_names", "slider_row_offsets", "slider_rows"}) if (in.count(k)) out[k] = in.at(k);
  return write_case(out_path, out) ? 0 : 1;
}

// ------------------------------------------------- skeleton export ("next" row f1)
// Joints in depth-first order (a parent always precedes its children), with what
// MovingPartBase::do_update / CharacterJoint::update_internals read per joint.
static void collect_joints(PartGroup *part, int parent_joint, std::vector<CharacterJoint *> &joints, std::vector<int32_t> &parents) {
  int me = parent_joint;
  if (part->is_character_joint()) {
    me = (int)joints.size();
    joints.push_back(DCAST(CharacterJoint, part));
    parents.push_back(parent_joint);
  }
  for (int i = 0; i < part->get_num_children(); ++i) collect_joints(part->get_child(i), me, joints, parents);
}

static void export_skeleton(Character *ch, const Exported &ex, Case &c) {
  PartBundle *bundle = ch->get_bundle(0);
  std::vector<CharacterJoint *> joints; std::vector<int32_t> parents;
  collect_joints(bundle, -1, joints, parents);
  int J = (int)joints.size();
  std::vector<int32_t> has_channel(J, 0), tlen(J * 12, 0), toff(J * 12, 0);
  std::vector<float> tables, defaults(J * 16), inverse(J * 16);
  for (int j = 0; j < J; ++j) {
    CharacterJoint *joint = joints[j];
    LMatrix4f dv = LCAST(float, joint->get_default_value());
    memcpy(&defaults[j * 16], dv.get_data(), 64);
    LMatrix4f inv = LCAST(float, joint->_initial_net_transform_inverse);
    memcpy(&inverse[j * 16], inv.get_data(), 64);
    AnimChannelBase *bound = joint->get_max_bound() > 0 ? joint->get_bound(0) : nullptr;
    if (bound != nullptr && bound->is_of_type(AnimChannelMatrixXfmTable::get_class_type())) {
      AnimChannelMatrixXfmTable *chan = DCAST(AnimChannelMatrixXfmTable, bound);
      has_channel[j] = 1;
      for (int i = 0; i < 12; ++i) {
        char id = "ijkabchprxyz"[i];  // matrix_component_letters (chan/animChannelMatrixXfmTable.cxx)
        CPTA_stdfloat t = chan->get_table(id);
        toff[j * 12 + i] = (int32_t)tables.size();
        tlen[j * 12 + i] = (int32_t)t.size();
        for (size_t k = 0; k < t.size(); ++k) tables.push_back((float)t[k]);
      }
    }
  }
  std::vector<int32_t> palette_joint;
  for (const VertexTransform *t : ex.palette) {
    int idx = -1;
    if (t->is_of_type(JointVertexTransform::get_class_type())) {
      const CharacterJoint *joint = ((const JointVertexTransform *)t)->get_joint();
      for (int j = 0; j < J; ++j) if (joints[j] == joint) idx = j;
    }
    palette_joint.push_back(idx);
  }
  // morph sliders driven by scalar channels: per slider of the SliderTable (the order the mesh's sliders have) its
  // CharacterSlider's bound AnimChannelScalarTable (char/characterVertexSlider.cxx:56-59, chan/animChannelScalarTable.cxx)
  std::vector<int32_t> s_has, s_len, s_off;
  std::vector<float> s_default;
  if (ex.slider_table != nullptr) {
    for (size_t si = 0; si < ex.slider_table->get_num_sliders(); ++si) {
      const VertexSlider *vs = ex.slider_table->get_slider(si);
      int has = 0, len = 0, off = (int)tables.size();
      float dflt = 0.0f;
      if (vs->is_of_type(CharacterVertexSlider::get_class_type())) {
        const CharacterSlider *cslider = ((const CharacterVertexSlider *)vs)->get_char_slider();
        dflt = (float)cslider->get_default_value();
        AnimChannelBase *bound = cslider->get_max_bound() > 0 ? cslider->get_bound(0) : nullptr;
        if (bound != nullptr && bound->is_of_type(AnimChannelScalarTable::get_class_type())) {
          CPTA_stdfloat t = DCAST(AnimChannelScalarTable, bound)->get_table();
          has = 1;
          len = (int)t.size();
          for (size_t k = 0; k < t.size(); ++k) tables.push_back((float)t[k]);
        }
      }
      s_has.push_back(has); s_len.push_back(len); s_off.push_back(off); s_default.push_back(dflt);
    }
  }
  c["skel_slider_has_channel"] = make_blob(DT_I32, s_has, {(uint64_t)s_has.size()});
  c["skel_slider_table_length"] = make_blob(DT_I32, s_len, {(uint64_t)s_len.size()});
  c["skel_slider_table_offset"] = make_blob(DT_I32, s_off, {(uint64_t)s_off.size()});
  c["skel_slider_default"] = make_blob(DT_F32, s_default, {(uint64_t)s_default.size()});
  LMatrix4f root = LCAST(float, bundle->get_root_xform());
  std::vector<float> rootv(root.get_data(), root.get_data() + 16);
  std::vector<int32_t> meta = {J, (int32_t)tables.size()};
  c["skel_meta"] = make_blob(DT_I32, meta, {2});
  c["skel_parent"] = make_blob(DT_I32, parents, {(uint64_t)J});
  c["skel_has_channel"] = make_blob(DT_I32, has_channel, {(uint64_t)J});
  c["skel_table_length"] = make_blob(DT_I32, tlen, {(uint64_t)J, 12});
  c["skel_table_offset"] = make_blob(DT_I32, toff, {(uint64_t)J, 12});
  c["skel_tables"] = make_blob(DT_F32, tables, {(uint64_t)tables.size()});
  c["skel_default_value"] = make_blob(DT_F32, defaults, {(uint64_t)J, 16});
  c["skel_initial_inverse"] = make_blob(DT_F32, inverse, {(uint64_t)J, 16});
  c["skel_root_xform"] = make_blob(DT_F32, rootv, {16});
  c["skel_palette_joint"] = make_blob(DT_I32, palette_joint, {(uint64_t)palette_joint.size()});
}

// The Actor of config C1: model + animation loaded through the reference's egg loader and bound with auto_bind.
struct Actor {
  PT(PandaNode) root;
  AnimControlCollection ctrls;
  Character *ch = nullptr;
  CPT(GeomVertexData) vd;
};

static bool load_actor(const char *model_path, const char *anim_path, Actor &a) {
  const bool one_file = strcmp(model_path, anim_path) == 0;  // e.g. models/ripple.egg: model and animation in one egg
  PT(PandaNode) model = load_egg_file(model_path);
  PT(PandaNode) anim = one_file ? nullptr : load_egg_file(anim_path);
  if (!model || (!one_file && !anim)) { fprintf(stderr, "egg load failed\n"); return false; }
  a.root = new PandaNode("root"); a.root->add_child(model);
  if (anim != nullptr) a.root->add_child(anim);
  auto_bind(a.root, a.ctrls, ~0);
  if (a.ctrls.get_num_anims() < 1) { fprintf(stderr, "no animation bound\n"); return false; }
  NodePath np(a.root);
  a.ch = DCAST(Character, np.find("**/+Character").node());
  NodePathCollection gns = np.find_all_matches("**/+GeomNode");
  for (int g = 0; g < gns.get_num_paths() && a.vd == nullptr; ++g) {
    GeomNode *gn = DCAST(GeomNode, gns.get_path(g).node());
    for (int i = 0; i < gn->get_num_geoms(); ++i) {
      CPT(GeomVertexData) cand = gn->get_geom(i)->get_vertex_data();
      if (cand->get_format()->get_animation().get_animation_type() == GeomEnums::AT_panda && cand->get_transform_blend_table() != nullptr) { a.vd = cand; break; }
    }
  }
  if (a.vd == nullptr) { fprintf(stderr, "no AT_panda vertex data\n"); return false; }
  return true;
}

static int cmd_egg(int argc, char **argv) {
  const char *out_path = argv[4];
  Actor actor;
  if (!load_actor(argv[2], argv[3], actor)) return 1;
  AnimControlCollection &ctrls = actor.ctrls;
  Character *ch = actor.ch;
  CPT(GeomVertexData) vd = actor.vd;
  Case out;
  Exported ex = export_vertex_data(vd, out);
  export_skeleton(ch, ex, out);
  // Frames with a fractional part switch on PartBundle::set_frame_blend_flag: the joints then take
  // MovingPartMatrix::get_blend_value's BT_linear branch between get_frame() and get_next_frame()
  // (chan/movingPartMatrix.cxx:78-113); what AnimControl reports is exported so the same triple can be replayed.
  std::vector<int32_t> frames, next_frames;
  std::vector<float> fracs;
  bool blend = false;
  for (int i = 5; i < argc; ++i) blend = blend || strchr(argv[i], '.') != nullptr;
  if (blend) ch->get_bundle(0)->set_frame_blend_flag(true);
  const char *bt = getenv("P3CS_BLEND_TYPE");   // "linear": PartBundle::BT_linear instead of the PRC default
  int blend_type = 1;
  if (bt != nullptr && strcmp(bt, "linear") == 0) {
    ch->get_bundle(0)->set_blend_type(PartBundle::BT_linear);
    blend_type = 0;
  } else if (bt != nullptr && strcmp(bt, "componentwise") == 0) {
    ch->get_bundle(0)->set_blend_type(PartBundle::BT_componentwise);
    blend_type = 2;
  } else if (bt != nullptr && strcmp(bt, "componentwise_quat") == 0) {
    ch->get_bundle(0)->set_blend_type(PartBundle::BT_componentwise_quat);
    blend_type = 3;
  }
  AnimControl *control = ctrls.get_anim(0);
  // Forced channels (MovingPartBase::_forced_channel, chan/movingPartMatrix.cxx:53-60): P3CS_FREEZE lists joints (in the
  // exported order) frozen with PartBundle::freeze_joint(name, pos, hpr, scale), P3CS_CONTROL joints driven by a node
  // through PartBundle::control_joint whose transform changes with every sample.  What each forced channel returns
  // (get_value(0, mat)) is exported per sample, so the same values can be replayed.
  std::vector<CharacterJoint *> all_joints; std::vector<int32_t> all_parents;
  collect_joints(ch->get_bundle(0), -1, all_joints, all_parents);
  std::vector<PT(PandaNode)> control_nodes;
  auto parse_list = [](const char *e) { std::vector<int> v; if (e) { const char *p = e; while (*p) { v.push_back(atoi(p)); while (*p && *p != ',') ++p; if (*p) ++p; } } return v; };
  for (int j : parse_list(getenv("P3CS_FREEZE"))) {
    if (j < 0 || j >= (int)all_joints.size()) { fprintf(stderr, "bad joint %d\n", j); return 1; }
    if (!ch->get_bundle(0)->freeze_joint(all_joints[j]->get_name(), LVecBase3(0.1f * j, -0.2f, 0.05f * j), LVecBase3(10.0f * j, -20.0f, 5.0f),
                                         LVecBase3(1.0f, 1.0f + 0.01f * j, 0.9f))) { fprintf(stderr, "freeze_joint failed\n"); return 1; }
  }
  for (int j : parse_list(getenv("P3CS_CONTROL"))) {
    if (j < 0 || j >= (int)all_joints.size()) { fprintf(stderr, "bad joint %d\n", j); return 1; }
    PT(PandaNode) node = new PandaNode("control");
    if (!ch->get_bundle(0)->control_joint(all_joints[j]->get_name(), node)) { fprintf(stderr, "control_joint failed\n"); return 1; }
    control_nodes.push_back(node);
  }
  std::vector<int32_t> forced_joints;
  for (size_t j = 0; j < all_joints.size(); ++j) if (all_joints[j]->get_forced_channel() != nullptr) forced_joints.push_back((int32_t)j);
  std::vector<float> forced_values;
  for (int i = 5; i < argc; ++i) {
    double fr = atof(argv[i]);
    for (size_t k = 0; k < control_nodes.size(); ++k) {
      const float t = (float)(i - 5) + 0.25f * (float)k;
      control_nodes[k]->set_transform(TransformState::make_pos_hpr_scale(LVecBase3(0.3f * t, 0.1f, -0.2f * t), LVecBase3(15.0f * t, 7.0f, -11.0f * t),
                                                                         LVecBase3(1.0f, 1.0f, 1.0f + 0.05f * t)));
    }
    control->pose(fr);
    ch->force_update();
    record_frame(vd, ex, out, (int)frames.size());
    frames.push_back(control->get_frame());
    next_frames.push_back(control->get_next_frame());
    fracs.push_back((float)(PN_stdfloat)control->get_frac());
    for (int32_t j : forced_joints) {
      LMatrix4 m;
      DCAST(AnimChannelMatrix, all_joints[j]->get_forced_channel())->get_value(0, m);
      LMatrix4f mf = LCAST(float, m);
      forced_values.insert(forced_values.end(), mf.get_data(), mf.get_data() + 16);
    }
  }
  if (!forced_joints.empty()) {
    out["forced_joints"] = make_blob(DT_I32, forced_joints, {(uint64_t)forced_joints.size()});
    out["forced_values"] = make_blob(DT_F32, forced_values, {(uint64_t)frames.size(), (uint64_t)forced_joints.size(), 16});
  }
  out["anim_frames"] = make_blob(DT_I32, frames, {(uint64_t)frames.size()});
  if (blend) {
    out["anim_next_frames"] = make_blob(DT_I32, next_frames, {(uint64_t)frames.size()});
    out["anim_fracs"] = make_blob(DT_F32, fracs, {(uint64_t)frames.size()});
    out["anim_blend_type"] = make_blob(DT_I32, std::vector<int32_t>{blend_type}, {1});
  }
  fprintf(stderr, "egg: rows=%d transforms=%d blends=%d frames=%d\n", vd->get_num_rows(), (int)ex.palette.size(),
          (int)vd->get_transform_blend_table()->get_num_blends(), (int)frames.size());
  return write_case(out_path, out) ? 0 : 1;
}

// Several AnimControls at once (PartBundle::set_anim_blend_flag + set_control_effect): the animation is bound a
// second time, every sample poses the two controls at their own frames with their own effects
// ("f0:e0,f1:e1" per sample; fractional frames switch the frame-blend flag on as in cmd_egg).  The per-control
// state is exported in the order PartBundle iterates its _blend map (ascending AnimControl address), which is the
// order MovingPartMatrix::get_blend_value sums in (chan/movingPartMatrix.cxx:82-198).
static int cmd_eggmulti(int argc, char **argv) {
  const char *out_path = argv[4];
  Actor actor;
  if (!load_actor(argv[2], argv[3], actor)) return 1;
  PartBundle *bundle = actor.ch->get_bundle(0);
  NodePath np(actor.root);
  AnimBundleNode *abn = DCAST(AnimBundleNode, np.find("**/+AnimBundleNode").node());
  std::vector<PT(AnimControl)> controls = {actor.ctrls.get_anim(0), bundle->bind_anim(abn->get_bundle(), ~0)};
  if (controls[1] == nullptr) { fprintf(stderr, "second bind failed\n"); return 1; }
  std::sort(controls.begin(), controls.end(), [](const PT(AnimControl) &a, const PT(AnimControl) &b) { return a.p() < b.p(); });
  bundle->set_anim_blend_flag(true);
  bool blend = false;
  for (int i = 5; i < argc; ++i) blend = blend || strchr(argv[i], '.') != nullptr;
  if (blend) bundle->set_frame_blend_flag(true);
  const char *bt = getenv("P3CS_BLEND_TYPE");
  int blend_type = 1;
  if (bt != nullptr && strcmp(bt, "linear") == 0) { bundle->set_blend_type(PartBundle::BT_linear); blend_type = 0; }
  else if (bt != nullptr && strcmp(bt, "componentwise") == 0) { bundle->set_blend_type(PartBundle::BT_componentwise); blend_type = 2; }
  else if (bt != nullptr && strcmp(bt, "componentwise_quat") == 0) { bundle->set_blend_type(PartBundle::BT_componentwise_quat); blend_type = 3; }
  Case out;
  Exported ex = export_vertex_data(actor.vd, out);
  export_skeleton(actor.ch, ex, out);
  std::vector<int32_t> frames, next_frames;
  std::vector<float> fracs, effects;
  int samples = 0;
  for (int i = 5; i < argc; ++i, ++samples) {
    double f[2] = {0, 0}, e[2] = {0, 0};
    if (sscanf(argv[i], "%lf:%lf,%lf:%lf", &f[0], &e[0], &f[1], &e[1]) != 4) { fprintf(stderr, "bad sample %s\n", argv[i]); return 1; }
    for (int k = 0; k < 2; ++k) {
      controls[k]->pose(f[k]);
      bundle->set_control_effect(controls[k], (PN_stdfloat)e[k]);
    }
    actor.ch->force_update();
    record_frame(actor.vd, ex, out, samples);
    for (int k = 0; k < 2; ++k) {
      frames.push_back(controls[k]->get_frame());
      next_frames.push_back(controls[k]->get_next_frame());
      fracs.push_back((float)(PN_stdfloat)controls[k]->get_frac());
      effects.push_back((float)bundle->get_control_effect(controls[k]));
    }
  }
  out["ctrl_frames"] = make_blob(DT_I32, frames, {(uint64_t)samples, 2});
  out["ctrl_next_frames"] = make_blob(DT_I32, next_frames, {(uint64_t)samples, 2});
  out["ctrl_fracs"] = make_blob(DT_F32, fracs, {(uint64_t)samples, 2});
  out["ctrl_effects"] = make_blob(DT_F32, effects, {(uint64_t)samples, 2});
  out["anim_blend_type"] = make_blob(DT_I32, std::vector<int32_t>{blend_type}, {1});
  out["anim_frame_blend"] = make_blob(DT_I32, std::vector<int32_t>{blend ? 1 : 0}, {1});
  return write_case(out_path, out) ? 0 : 1;
}

// Times the reference CPU path: every iteration re-sets all matrices / sliders
// (forcing a full recompute) and calls animate_vertices(true, thread).
static int cmd_bench(const char *in_path, int iters, int warmup, int copies) {
  Case in; if (!read_case(in_path, in)) return 1;
  apply_recipe_coordinate_system(in);
  std::vector<Built> chars(copies);
  for (int i = 0; i < copies; ++i) if (!build_from_recipe(in, chars[i])) return 1;
  int frames = (int)in.at("palette").shape[0];
  int num_rows = in.at("meta").i32()[0];
  Thread *th = Thread::get_current_thread();
  double total = 0, best = 1e30;
  for (int it = 0; it < warmup + iters; ++it) {
    for (int i = 0; i < copies; ++i) set_frame(chars[i], in, (it + i) % frames);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < copies; ++i) { CPT(GeomVertexData) out = chars[i].vd->animate_vertices(true, th); (void)out; }
    auto t1 = std::chrono::steady_clock::now();
    double s = std::chrono::duration<double>(t1 - t0).count();
    if (it >= warmup) { total += s; best = std::min(best, s); }
  }
  double verts = (double)num_rows * copies;
  printf("{\"rows\": %d, \"characters\": %d, \"iters\": %d, \"mean_s\": %.9g, \"best_s\": %.9g, \"verts_per_s_mean\": %.6g, \"verts_per_s_best\": %.6g}\n",
         num_rows, copies, iters, total / iters, best, verts / (total / iters), verts / best);
  return 0;
}

// ------------------------------------------------- AT_hardware conversion ("next" row f3)
// Runs the reference's own Panda-table -> hardware-table conversion (GeomVertexData::copy_from,
// gobj/geomVertexData.cxx:574-648, indexed branch) through convert_to() and exports what it wrote: the
// transform_weight / transform_index columns and the TransformTable order.
static int cmd_hw(const char *in_path, const char *out_path) {
  Case in; if (!read_case(in_path, in)) return 1;
  Built b; if (!build_from_recipe(in, b)) return 1;
  int num_rows = in.at("meta").i32()[0];
  PT(GeomVertexFormat) f = new GeomVertexFormat(*b.vd->get_format());
  f->remove_column(InternalName::get_transform_blend());
  PT(GeomVertexArrayFormat) af = new GeomVertexArrayFormat;
  af->add_column(InternalName::get_transform_weight(), 4, GeomEnums::NT_float32, GeomEnums::C_other);
  af->add_column(InternalName::get_transform_index(), 4, GeomEnums::NT_uint16, GeomEnums::C_index);
  f->add_array(af);
  GeomVertexAnimationSpec spec; spec.set_hardware(4, true); f->set_animation(spec);
  CPT(GeomVertexFormat) fmt = GeomVertexFormat::register_format(f);
  CPT(GeomVertexData) hw = b.vd->convert_to(fmt);
  if (hw->get_transform_table() == nullptr) { fprintf(stderr, "no TransformTable after convert_to\n"); return 1; }
  std::vector<float> weights((size_t)num_rows * 4);
  std::vector<int32_t> indices((size_t)num_rows * 4);
  GeomVertexReader rw(hw, InternalName::get_transform_weight()), ri(hw, InternalName::get_transform_index());
  for (int r = 0; r < num_rows; ++r) {
    LVecBase4 w = rw.get_data4();
    LVecBase4i i = ri.get_data4i();
    for (int k = 0; k < 4; ++k) { weights[(size_t)r * 4 + k] = (float)w[k]; indices[(size_t)r * 4 + k] = i[k]; }
  }
  const TransformTable *tt = hw->get_transform_table();
  std::vector<int32_t> table;
  for (size_t t = 0; t < tt->get_num_transforms(); ++t) {
    int found = -1;
    for (size_t j = 0; j < b.transforms.size(); ++j) if (b.transforms[j].p() == tt->get_transform(t)) found = (int)j;
    table.push_back(found);
  }
  Case out;
  out["hw_weight"] = make_blob(DT_F32, weights, {(uint64_t)num_rows, 4});
  out["hw_index"] = make_blob(DT_I32, indices, {(uint64_t)num_rows, 4});
  out["hw_table"] = make_blob(DT_I32, table, {(uint64_t)table.size()});
  return write_case(out_path, out) ? 0 : 1;
}

int main(int argc, char **argv) {
  if (const char *prc = getenv("P3CS_PRC")) load_prc_file_data("harness-extra", prc);   // e.g. "vertex-colors-prefer-packed 1"
  if (argc >= 4 && strcmp(argv[1], "hw") == 0) return cmd_hw(argv[2], argv[3]);
  if (argc >= 4 && strcmp(argv[1], "run") == 0) return cmd_run(argv[2], argv[3]);
  if (argc >= 6 && strcmp(argv[1], "egg") == 0) return cmd_egg(argc, argv);
  if (argc >= 6 && strcmp(argv[1], "eggmulti") == 0) return cmd_eggmulti(argc, argv);
  if (argc >= 5 && strcmp(argv[1], "bench") == 0) return cmd_bench(argv[2], atoi(argv[3]), atoi(argv[4]), argc > 5 ? atoi(argv[5]) : 1);
  fprintf(stderr, "usage: ref_harness run <recipe> <golden> | hw <recipe> <out> | egg <model> <anim> <golden> <frame>... | bench <recipe> <iters> <warmup> [characters]\n");
  return 2;
}
