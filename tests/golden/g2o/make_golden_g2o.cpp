// Recipe that PINS the g2o / Sophus / Eigen half of the oracle where those libraries exist (they do not in the build image:
// DESIGN.md §3 "parity unpinned").  It feeds the committed problems of tests/golden/g2o/g2o_problems.txt (written by
// export_problems.py from golden_geom.npz + one pose graph) through the REFERENCE's own vertex / edge classes
// (include/StereoVisionSLAM/g2o_types.h, compiled from the reference checkout — nothing is copied into this repo) with
// exactly the solver set-ups of the reference:
//   pose-only   Frontend::EstimateCurrentPose        src/frontend.cpp:408-527   (BlockSolver_6_3 + dense, LM, 4 x optimize(10))
//   window BA   Backend::Optimize                    src/backend.cpp:22-164     (Schur, dense, Huber(chi2_th), optimize(10))
//   pose graph  LoopClosure::PoseGraphOptimization   src/loopclosure.cpp:641-746 (BlockSolver<6,6> + dense, optimize(22))
// and writes golden_g2o.txt; import_results.py turns it into tests/golden/golden_g2o.npz, which tests/test_golden.py
// compares the C oracle and the CUDA path against when present.  Build: see CMakeLists.txt beside this file.
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include "StereoVisionSLAM/g2o_types.h"

using namespace slam;
typedef Eigen::Matrix<double, 7, 1> Vec7;

static Sophus::SE3d toSE3(const double *p)      // [qx qy qz qw tx ty tz]
{
    return Sophus::SE3d(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
}
static void put(std::ostream &o, const Sophus::SE3d &T)
{
    const Eigen::Quaterniond q = T.unit_quaternion();
    o << q.x() << ' ' << q.y() << ' ' << q.z() << ' ' << q.w() << ' ' << T.translation().transpose() << '\n';
}
static g2o::OptimizationAlgorithmLevenberg *make_63()
{
    typedef g2o::BlockSolver_6_3 BS;
    typedef g2o::LinearSolverDense<BS::PoseMatrixType> LS;
    return new g2o::OptimizationAlgorithmLevenberg(std::make_unique<BS>(std::make_unique<LS>()));
}

int main(int argc, char **argv)
{
    std::ifstream in(argc > 1 ? argv[1] : "g2o_problems.txt");
    std::ofstream out(argc > 2 ? argv[2] : "golden_g2o.txt");
    out << std::setprecision(17);
    std::string tag;
    while (in >> tag) {
        if (tag == "POSE_ONLY") {      // m, K(4), T0(7), then m x (pw xyz, uv)
            int m; double K4[4], T0[7];
            in >> m; for (double &v : K4) in >> v; for (double &v : T0) in >> v;
            Eigen::Matrix3d K; K << K4[0], 0, K4[2], 0, K4[1], K4[3], 0, 0, 1;
            g2o::SparseOptimizer opt; opt.setAlgorithm(make_63());
            VertexPose *vp = new VertexPose(); vp->setId(0); vp->setEstimate(toSE3(T0)); opt.addVertex(vp);
            std::vector<EdgeProjectionPoseOnly *> edges; std::vector<char> outlier(m, 0);
            for (int i = 0; i < m; i++) {
                Eigen::Vector3d pw; Eigen::Vector2d uv; in >> pw[0] >> pw[1] >> pw[2] >> uv[0] >> uv[1];
                EdgeProjectionPoseOnly *e = new EdgeProjectionPoseOnly(pw, K);
                e->setId(i + 1); e->setVertex(0, vp); e->setMeasurement(uv); e->setInformation(Eigen::Matrix2d::Identity());
                e->setRobustKernel(new g2o::RobustKernelHuber); edges.push_back(e); opt.addEdge(e);
            }
            for (int it = 0; it < 4; it++) {
                vp->setEstimate(toSE3(T0)); opt.initializeOptimization(); opt.optimize(10);
                for (int i = 0; i < m; i++) {
                    if (outlier[i]) edges[i]->computeError();
                    if (edges[i]->chi2() > 5.991) { outlier[i] = 1; edges[i]->setLevel(1); } else { outlier[i] = 0; edges[i]->setLevel(0); }
                    if (it == 2) edges[i]->setRobustKernel(nullptr);
                }
            }
            out << "POSE_ONLY " << m << '\n'; put(out, vp->estimate());
            for (int i = 0; i < m; i++) out << int(outlier[i]) << (i + 1 < m ? ' ' : '\n');
        } else if (tag == "BA") {      // N L E, Kl(4) Kr(4) extl(7) extr(7) huber, N poses, L points, E x (kf lm cam u v)
            int N, L, E; double Kl[4], Kr[4], el[7], er[7], huber;
            in >> N >> L >> E; for (double &v : Kl) in >> v; for (double &v : Kr) in >> v; for (double &v : el) in >> v; for (double &v : er) in >> v; in >> huber;
            Eigen::Matrix3d K[2]; K[0] << Kl[0], 0, Kl[2], 0, Kl[1], Kl[3], 0, 0, 1; K[1] << Kr[0], 0, Kr[2], 0, Kr[1], Kr[3], 0, 0, 1;
            Sophus::SE3d ext[2] = {toSE3(el), toSE3(er)};
            g2o::SparseOptimizer opt; opt.setAlgorithm(make_63());
            std::vector<VertexPose *> vp(N); std::vector<VertexXYZ *> vl(L, nullptr); std::vector<Eigen::Vector3d> pts(L);
            for (int i = 0; i < N; i++) { double p[7]; for (double &v : p) in >> v; vp[i] = new VertexPose(); vp[i]->setId(i); vp[i]->setEstimate(toSE3(p)); opt.addVertex(vp[i]); }
            for (int i = 0; i < L; i++) in >> pts[i][0] >> pts[i][1] >> pts[i][2];
            std::vector<EdgeProjection *> edges(E);
            for (int e = 0; e < E; e++) {
                int kf, lm, cam; Eigen::Vector2d uv; in >> kf >> lm >> cam >> uv[0] >> uv[1];
                if (!vl[lm]) { vl[lm] = new VertexXYZ(); vl[lm]->setEstimate(pts[lm]); vl[lm]->setId(lm + N); vl[lm]->setMarginalized(true); opt.addVertex(vl[lm]); }
                EdgeProjection *ed = new EdgeProjection(K[cam], ext[cam]);
                ed->setId(e + 1); ed->setVertex(0, vp[kf]); ed->setVertex(1, vl[lm]); ed->setMeasurement(uv); ed->setInformation(Eigen::Matrix2d::Identity());
                auto rk = new g2o::RobustKernelHuber(); rk->setDelta(huber); ed->setRobustKernel(rk); opt.addEdge(ed); edges[e] = ed;
            }
            opt.initializeOptimization(); opt.optimize(10);
            out << "BA " << N << ' ' << L << ' ' << E << '\n';
            for (int i = 0; i < N; i++) put(out, vp[i]->estimate());
            for (int i = 0; i < L; i++) out << (vl[i] ? vl[i]->estimate() : pts[i]).transpose() << '\n';
            for (int e = 0; e < E; e++) out << edges[e]->chi2() << (e + 1 < E ? ' ' : '\n');
        } else if (tag == "POSE_GRAPH") {      // N E, N x (fixed, pose), E x (a b meas)
            int N, E; in >> N >> E;
            typedef g2o::BlockSolver<g2o::BlockSolverTraits<6, 6>> BS;
            typedef g2o::LinearSolverDense<BS::PoseMatrixType> LS;
            g2o::SparseOptimizer opt;
            opt.setAlgorithm(new g2o::OptimizationAlgorithmLevenberg(std::make_unique<BS>(std::make_unique<LS>())));
            std::vector<VertexPose *> vp(N);
            for (int i = 0; i < N; i++) { int fx; double p[7]; in >> fx; for (double &v : p) in >> v; vp[i] = new VertexPose(); vp[i]->setId(i); vp[i]->setEstimate(toSE3(p)); vp[i]->setMarginalized(false); if (fx) vp[i]->setFixed(true); opt.addVertex(vp[i]); }
            for (int e = 0; e < E; e++) {
                int a, b; double p[7]; in >> a >> b; for (double &v : p) in >> v;
                EdgePoseGraph *ed = new EdgePoseGraph(); ed->setId(e); ed->setVertex(0, vp[a]); ed->setVertex(1, vp[b]); ed->setMeasurement(toSE3(p));
                ed->setInformation(Eigen::Matrix<double, 6, 6>::Identity()); opt.addEdge(ed);
            }
            opt.initializeOptimization(); opt.optimize(22);
            out << "POSE_GRAPH " << N << '\n';
            for (int i = 0; i < N; i++) put(out, vp[i]->estimate());
        }
    }
    std::cout << "wrote golden_g2o.txt" << std::endl;
    return 0;
}
