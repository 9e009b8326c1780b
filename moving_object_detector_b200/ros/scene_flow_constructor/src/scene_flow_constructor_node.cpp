// scene_flow_constructor: the node executable, same name as the reference's
// (scene_flow_constructor/src/scene_flow_constructor_node.cpp:3-10, CMakeLists.txt:43-46).
#include <ros/ros.h>

#include "scene_flow_constructor_b200_node.h"

int main(int argc, char** argv) {
  ros::init(argc, argv, "scene_flow_constructor");
  scene_flow_constructor::SceneFlowConstructor node;
  ros::spin();
  return 0;
}
