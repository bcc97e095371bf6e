// stand-in for <ros/ros.h>: src/ORBextractor.cc includes it but uses nothing from it
#define ROS_ASSERT(x) ((void)0)
