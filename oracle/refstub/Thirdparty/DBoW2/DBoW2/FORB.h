// TEST INFRASTRUCTURE: RGBDFrame carries a DBoW2::BowVector member (include/rgbdframe.h:55); the mapper path never reads it.
#ifndef SSM_REFSTUB_DBOW2
#define SSM_REFSTUB_DBOW2
namespace DBoW2 { struct BowVector {}; }
#endif
