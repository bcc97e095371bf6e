// stand-in for <ros/ros.h>: the compiled reference files include it for logging / assertions only
#include <iostream>
#include <cstdlib>
#include <cmath>
#define ROS_ASSERT(x) ((void)0)
