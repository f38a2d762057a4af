#include <pcl/visualization/cloud_viewer.h>
