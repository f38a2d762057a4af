// TEST INFRASTRUCTURE: pcl::VoxelGrid<PointXYZRGBA> -- the WRITTEN definition of SURVEY App. B-2 (fp32 inverse leaf,
// floor(coord * inv) cells relative to the bounding box, sequential fp32 sums per cell in index order, truncated mean
// colour, output in ascending cell index).  Not the real library: PCL's semantics stay unpinned.
#ifndef SSM_REFSTUB_PCL_VOXEL_GRID
#define SSM_REFSTUB_PCL_VOXEL_GRID
#include <algorithm>
#include <cmath>
#include <pcl/point_types.h>
namespace pcl {
template <typename PointT> class VoxelGrid {
public:
    void setLeafSize(float lx, float ly, float lz) { inv_[0] = 1.0f / lx; inv_[1] = 1.0f / ly; inv_[2] = 1.0f / lz; }
    void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in_ = c; }
    void filter(PointCloud<PointT>& out)
    {
        out.clear();
        out.is_dense = true;
        struct Item { long long idx; size_t i; };
        std::vector<Item> items;
        long long mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
        bool first = true;
        for (const PointT& p : in_->points) {
            if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
            const long long c[3] = {(long long)std::floor(p.x * inv_[0]), (long long)std::floor(p.y * inv_[1]), (long long)std::floor(p.z * inv_[2])};
            for (int a = 0; a < 3; ++a) {
                if (first || c[a] < mn[a]) mn[a] = c[a];
                if (first || c[a] > mx[a]) mx[a] = c[a];
            }
            first = false;
        }
        const long long dx = mx[0] - mn[0] + 1, dy = mx[1] - mn[1] + 1;
        for (size_t i = 0; i < in_->points.size(); ++i) {
            const PointT& p = in_->points[i];
            if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
            const long long ci = (long long)std::floor(p.x * inv_[0]) - mn[0], cj = (long long)std::floor(p.y * inv_[1]) - mn[1],
                            ck = (long long)std::floor(p.z * inv_[2]) - mn[2];
            items.push_back(Item{ci + cj * dx + ck * dx * dy, i});
        }
        std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.idx < b.idx; });
        for (size_t s = 0; s < items.size();) {
            size_t e = s;
            float sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0;
            while (e < items.size() && items[e].idx == items[s].idx) {
                const PointT& p = in_->points[items[e].i];
                sx += p.x; sy += p.y; sz += p.z; sr += (float)p.r; sg += (float)p.g; sb += (float)p.b;
                ++e;
            }
            const float n = (float)(e - s);
            PointT q;
            q.x = sx / n; q.y = sy / n; q.z = sz / n;
            q.rgba = ((uint32_t)(int)(sr / n) << 16) | ((uint32_t)(int)(sg / n) << 8) | (uint32_t)(int)(sb / n);
            out.points.push_back(q);
            s = e;
        }
        out.width = (uint32_t)out.points.size(); out.height = 1;
    }
private:
    float inv_[3] = {1, 1, 1};
    typename PointCloud<PointT>::Ptr in_;
};
}  // namespace pcl
#endif
