"""TEST INFRASTRUCTURE. Cuts the reference's own CALLER code for the drop-in boundary out of /root/reference (never
copied into the repository: the output goes to the git-ignored oracle/_ref/) so that it can be compiled against both
the reference's class and the facade's:

  ros_utils.hpp   ProcessingStage + PointCloud2Iterators                             (hpp of src/ros/ros_utils.cpp)
  ros_utils.cpp   clusterToPointCloud, columnToPointCloud                            ros_utils.cpp:11-77
                  prepareMessageAndCreateIterators, addPointToMessage                ros_utils.cpp:108-298
  kitti_demo.cpp  KittiDemo::addColumnAndEvaluateFrameIfCompleted (callback body)    kitti_demo.cpp:173-224
                  KittiDemo::makePseudoFiringFromRangeImageColumn                    kitti_demo.cpp:123-159
and the functions of the rows SURVEY 8f-2 / 8f-4 name (oracle/cc_eval_driver.cpp compiles them against the reference's
own headers):
  kitti_evaluation.cpp  KittiEvaluation ctor, evaluate, evaluateGroundPoints, evaluateClusters, addPointForKey  :10-157
  kitti_loader.cpp      recoverLaserIndices, generateRangeImage, undoEgoMotionCorrection     kitti_loader.cpp:47-210
                        interpolate                                                          kitti_loader.cpp:297-328
                        getSemanticKittiLabel*Mapping                                        kitti_loader.cpp:566-613

Ranges are found by their first lines, not by line numbers.   usage: extract_caller_excerpts.py REFERENCE OUTDIR
"""
import os
import sys


def cut(path, start, stop):
    lines = open(path).read().split("\n")
    a = next(i for i, l in enumerate(lines) if l.startswith(start))
    b = next(i for i in range(a + 1, len(lines)) if lines[i].startswith(stop))
    return "\n".join(lines[a:b]) + "\n"


def main(ref, out):
    os.makedirs(out, exist_ok=True)
    hpp = os.path.join(ref, "include/continuous_clustering/ros/ros_utils.hpp")
    cpp = os.path.join(ref, "src/ros/ros_utils.cpp")
    demo = os.path.join(ref, "src/tools/kitti_demo.cpp")
    parts = {
        "ros_utils_types.inc": cut(hpp, "enum ProcessingStage", "sensor_msgs::PointCloud2Ptr clusterToPointCloud("),
        "ros_utils_clouds.inc": cut(cpp, "sensor_msgs::PointCloud2Ptr clusterToPointCloud(",
                                    "sensor_msgs::PointCloud2Ptr firingToPointCloud("),
        "ros_utils_fields.inc": cut(cpp, "PointCloud2Iterators prepareMessageAndCreateIterators(",
                                    "void addRawPointToMessage("),
        "kitti_demo_callback.inc": cut(demo, "    void addColumnAndEvaluateFrameIfCompleted(", "  public:"),
        "kitti_demo_pseudo_firing.inc": cut(demo, "    static inline RawPoints::Ptr makePseudoFiringFromRangeImageColumn(",
                                            "    void evaluatePreviousFrame()"),
        "kitti_eval_metrics.inc": cut(os.path.join(ref, "src/evaluation/kitti_evaluation.cpp"), "KittiEvaluation::KittiEvaluation()",
                                      "std::string KittiEvaluation::generateEvaluationResults()"),
        "kitti_loader_range_image.inc": cut(os.path.join(ref, "src/evaluation/kitti_loader.cpp"),
                                            "void KittiLoader::recoverLaserIndices(", "Oxts KittiLoader::loadSingleOxfordMeasurement("),
        "kitti_loader_interpolate.inc": cut(os.path.join(ref, "src/evaluation/kitti_loader.cpp"),
                                            "StampedPose KittiLoader::interpolate(", "std::vector<StampedPose> KittiLoader::getAllDynamicTransforms("),
        "kitti_loader_labels.inc": cut(os.path.join(ref, "src/evaluation/kitti_loader.cpp"),
                                       "std::map<uint16_t, std::string> KittiLoader::getSemanticKittiLabelNumericToLabelNameMapping()",
                                       "std::vector<std::string> KittiLoader::split("),
    }
    for name, text in parts.items():
        with open(os.path.join(out, name), "w") as f:
            f.write("// generated from the reference by oracle/extract_caller_excerpts.py -- do not commit\n" + text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
