// Link-time stand-ins for the symbols owned by the reference's GUI translation units
// (gl.cc, model_draw.cc, option.cc, vis_strip.cc, worldgui.cc, typetable.cc, canvas.cc),
// which the headless oracle build leaves out. Everything here is a no-op except the model
// type table. Test infrastructure only; never linked into the product library.
#include "stage.hh"
#include "option.hh"
using namespace Stg;

// --- model_draw.cc ---
void Model::DrawBlocks() {}
void Model::DrawStatus(Camera *) {}
void Model::DrawSelected() {}
void Model::DrawPicker() {}
void Model::DataVisualize(Camera *) {}
void Model::AddVisualizer(Visualizer *, bool) {}
void Model::RemoveVisualizer(Visualizer *) {}

// --- gl.cc ---
void Gl::pose_shift(const Pose &) {}
void Gl::pose_inverse_shift(const Pose &) {}
void Gl::coord_shift(double, double, double, double) {}
void Gl::draw_string(float, float, float, const char *) {}
void Gl::draw_centered_rect(float, float, float, float) {}
void Gl::draw_origin(double) {}

// --- option.cc ---
Option::Option(const std::string &n, const std::string &tok, const std::string &key, bool v, World *w)
    : optName(n), value(v), wf_token(tok), shortcut(key), menu(NULL), menuIndex(0), menuCb(NULL),
      menuCbWidget(NULL), _world(w), htname(n)
{
}

// --- vis_strip.cc ---
StripPlotVis::StripPlotVis(float x, float y, float w, float h, size_t len, Color fg, Color bg,
                           const char *name, const char *wfname)
    : Visualizer(name, wfname), data(NULL), len(len), count(0), x(x), y(y), w(w), h(h), min(0),
      max(0), fgcolor(fg), bgcolor(bg)
{
}
StripPlotVis::~StripPlotVis() {}
void StripPlotVis::Visualize(Model *, Camera *) {}
void StripPlotVis::AppendValue(float) {}

// --- worldgui.cc: model.cc does dynamic_cast<WorldGui*>, so the typeinfo/vtable must exist ---
WorldGui::WorldGui(int w, int h, const char *c) : World(), Fl_Window(w, h, c) {}
WorldGui::~WorldGui() {}
void WorldGui::Redraw() {}
std::string WorldGui::ClockString() const { return ""; }
bool WorldGui::Update() { return World::Update(); }
bool WorldGui::Load(const std::string &p) { return World::Load(p); }
bool WorldGui::Load(std::istream &s, const std::string &p) { return World::Load(s, p); }
void WorldGui::UnLoad() {}
bool WorldGui::Save(const char *f) { return World::Save(f); }
Model *WorldGui::RecentlySelectedModel() const { return NULL; }
void WorldGui::Start() {}
void WorldGui::Stop() {}
void WorldGui::RemoveChild(Model *m) { World::RemoveChild(m); }
void WorldGui::AddModel(Model *m) { World::AddModel(m); }
void WorldGui::PushColor(Color) {}
void WorldGui::PushColor(double, double, double, double) {}
void WorldGui::PopColor() {}

// --- typetable.cc, minus ModelCamera (OpenGL z-buffer sensor; no raycasting) ---
template <class T> static Model *make(World *w, Model *p, const std::string &t) { return new T(w, p, t); }
void Stg::RegisterModels()
{
  Model::name_map["model"] = make<Model>;
  Model::name_map["position"] = make<ModelPosition>;
  Model::name_map["ranger"] = make<ModelRanger>;
  Model::name_map["fiducial"] = make<ModelFiducial>;
  Model::name_map["blobfinder"] = make<ModelBlobfinder>;
  Model::name_map["bumper"] = make<ModelBumper>;
  Model::name_map["gripper"] = make<ModelGripper>;
  Model::name_map["actuator"] = make<ModelActuator>;
  Model::name_map["blinkenlight"] = make<ModelBlinkenlight>;
  Model::name_map["lightindicator"] = make<ModelLightIndicator>;
}
