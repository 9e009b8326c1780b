// scene_flow_clusterer_node: runs the clusterer nodelet in a process of its own, like the reference's wrapper
// executable of the same name (scene_flow_clusterer/src/clusterer_node.cpp:8-15).
#include <nodelet/loader.h>
#include <ros/ros.h>

int main(int argc, char** argv) {
  ros::init(argc, argv, "scene_flow_clusterer");
  nodelet::Loader loader(false);  // no load / unload services, as in the reference's wrapper (:8)
  const nodelet::M_string remappings;  // empty: the reference forwards no remappings (:10)
  const nodelet::V_string arguments(argv + 1, argv + argc);
  loader.load(ros::this_node::getName(), "scene_flow_clusterer/scene_flow_clusterer", remappings, arguments);
  ros::spin();
  return 0;
}
