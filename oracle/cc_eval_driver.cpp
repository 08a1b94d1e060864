// cc_eval_driver.cpp -- TEST INFRASTRUCTURE. The reference's own evaluation metrics and KITTI replay front-end, cut out
// of /root/reference at build time (oracle/extract_caller_excerpts.py -> oracle/_ref/*.inc, never committed) and compiled
// UNMODIFIED against the reference's own headers (Eigen / PCL stand-ins for what the image lacks), behind a small C API:
//   ev_evaluate        KittiEvaluation::evaluateGroundPoints + evaluateClusters      kitti_evaluation.cpp:44-146
//   ev_frame_to_firings  KittiLoader::recoverLaserIndices, undoEgoMotionCorrection, generateRangeImage,
//                      KittiDemo::makePseudoFiringFromRangeImageColumn, KittiLoader::interpolate -- the per-frame body
//                      of kitti_demo.cpp:369-403
// The oracle of SURVEY rows 8f-2 and 8f-4; nothing here is on the product path.
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>

#include <continuous_clustering/clustering/point_types.hpp>
#include <continuous_clustering/evaluation/kitti_evaluation.hpp>

namespace continuous_clustering
{
KittiLoader::KittiLoader() = default; // kitti_loader.cpp:10
#include "kitti_eval_metrics.inc"
#include "kitti_loader_range_image.inc"
#include "kitti_loader_interpolate.inc"
#include "kitti_loader_labels.inc"

struct KittiDemoPseudoFiring
{
#include "kitti_demo_pseudo_firing.inc"
    static RawPoints::Ptr make(const std::vector<KittiPoint>& range_image, uint64_t a, uint64_t b, int col, int seq, int frame)
    {
        return makePseudoFiringFromRangeImageColumn(range_image, a, b, col, seq, frame);
    }
};
} // namespace continuous_clustering

using namespace continuous_clustering;

#define EV_API extern "C" __attribute__((visibility("default")))

EV_API void ev_evaluate(int n, const uint16_t* semantic_label, const uint8_t* is_ground, const uint32_t* gt_label,
                        const uint32_t* det_label, double* out6)
{
    std::vector<KittiSegmentationEvaluationPoint> pc(static_cast<size_t>(n));
    for (int i = 0; i < n; i++)
    {
        pc[i].point.semantic_label = semantic_label[i];
        pc[i].is_ground_point = is_ground[i] != 0;
        pc[i].euclidean_clustering_label = gt_label[i];
        pc[i].detection_label = det_label[i];
    }
    KittiEvaluation ev;
    EvaluationResultForFrame r{};
    ev.evaluateGroundPoints(pc, r);
    KittiEvaluation::evaluateClusters(pc, r);
    out6[0] = r.tp;
    out6[1] = r.fn;
    out6[2] = r.fp;
    out6[3] = r.tn;
    out6[4] = r.over_segmentation_entropy;
    out6[5] = r.under_segmentation_entropy;
}

// One KITTI frame -> RANGE_IMAGE_WIDTH pseudo firings (kitti_demo.cpp:369-403). xyzi: n points x 4 floats in file order;
// pose_stamps / poses12: the sequence's odom_from_velodyne transforms (n_poses x 12 doubles, 3x4 row major) and their
// stamps; frame_pose12: the transform at the middle of this frame's rotation. Outputs: firings
// [RANGE_IMAGE_WIDTH][RANGE_IMAGE_HEIGHT] RawPoint records (48 bytes each), firing_poses [RANGE_IMAGE_WIDTH][12],
// laser_index[n] and range-image cell [n] (row * width + column, -1 if the point lost its cell) for intermediate checks.
EV_API int ev_frame_to_firings(int n, const float* xyzi, uint64_t stamp_start, uint64_t stamp_end, const double* frame_pose12,
                               int n_poses, const uint64_t* pose_stamps, const double* poses12, int sequence_index, int frame_index,
                               void* firings, double* firing_poses, uint8_t* laser_index, int32_t* cell_of_point, float* uncorrected_xyz)
{
    try
    {
        std::vector<KittiPoint> points(static_cast<size_t>(n));
        for (int i = 0; i < n; i++)
        {
            points[i].x = xyzi[4 * i];
            points[i].y = xyzi[4 * i + 1];
            points[i].z = xyzi[4 * i + 2];
            points[i].i = xyzi[4 * i + 3];
        }
        auto toIso = [](const double* m)
        {
            Eigen::Isometry3d t = Eigen::Isometry3d::Identity();
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 4; j++)
                    t(i, j) = m[i * 4 + j];
            return t;
        };
        std::vector<StampedPose> transforms(static_cast<size_t>(n_poses));
        for (int k = 0; k < n_poses; k++)
        {
            transforms[k].stamp = pose_stamps[k];
            transforms[k].pose = toIso(poses12 + 12 * k);
        }
        KittiLoader loader;
        loader.recoverLaserIndices(points);
        for (int i = 0; i < n; i++)
            laser_index[i] = points[i].laser_index;
        loader.undoEgoMotionCorrection(points, stamp_start, stamp_end, toIso(frame_pose12), transforms);
        for (int i = 0; i < n; i++)
        {
            uncorrected_xyz[3 * i] = points[i].x;
            uncorrected_xyz[3 * i + 1] = points[i].y;
            uncorrected_xyz[3 * i + 2] = points[i].z;
        }
        std::vector<KittiPoint> range_image = loader.generateRangeImage(points);
        for (int i = 0; i < n; i++)
            cell_of_point[i] = -1;
        for (size_t c = 0; c < range_image.size(); c++)
            if (range_image[c].original_kitti_index >= 0)
                cell_of_point[range_image[c].original_kitti_index] = static_cast<int32_t>(c);
        RawPoint* out = static_cast<RawPoint*>(firings);
        for (int col = 0; col < KittiLoader::RANGE_IMAGE_WIDTH; col++)
        {
            RawPoints::Ptr f = KittiDemoPseudoFiring::make(range_image, stamp_start, stamp_end, col, sequence_index, frame_index);
            std::memcpy(static_cast<void*>(out + static_cast<size_t>(col) * KittiLoader::RANGE_IMAGE_HEIGHT), f->points.data(),
                        sizeof(RawPoint) * KittiLoader::RANGE_IMAGE_HEIGHT);
            const Eigen::Isometry3d pose = loader.interpolate(transforms, f->stamp).pose;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 4; j++)
                    firing_poses[12 * col + i * 4 + j] = pose(i, j);
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        std::cerr << "ev_frame_to_firings: " << e.what() << std::endl;
        return 1;
    }
}

EV_API int ev_range_image_width(void)
{
    return KittiLoader::RANGE_IMAGE_WIDTH;
}
EV_API int ev_range_image_height(void)
{
    return KittiLoader::RANGE_IMAGE_HEIGHT;
}
