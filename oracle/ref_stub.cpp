// oracle/ref_stub.cpp -- TEST INFRASTRUCTURE.  Minimal pybind module around the reference's
// own prim3d::marching_cubes and prim3d::save_mesh_as_ply (declared in
// /root/reference/src/prim3d/Utility/marching_cubes.h:14-17, both defined in marching_cubes.cu),
// so the unmodified reference kernels can be run beside ours on the GPU box.  The reference's full
// bindings.cpp also pulls in the OptiX/Eigen ray caster, which is out of scope and not buildable here.
#include <torch/extension.h>

#include <string>
#include <vector>

namespace prim3d {
std::vector<torch::Tensor> marching_cubes(const torch::Tensor &density_grid, const float thresh,
                                          const std::vector<float> lower, const std::vector<float> upper);
void save_mesh_as_ply(const std::string filename, torch::Tensor vertices, torch::Tensor faces, torch::Tensor colors);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "reference prim3d::marching_cubes compiled unmodified for sm_100a (oracle tier T1)";
    m.def("marching_cubes", &prim3d::marching_cubes);
    m.def("save_mesh_as_ply", &prim3d::save_mesh_as_ply);
}
