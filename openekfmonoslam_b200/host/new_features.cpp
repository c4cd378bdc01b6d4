// new_features.cpp -- host side of the new-feature policy: which of the frame's keypoints become map features
// (kalmanFilter/modules/1PointRansacEKF/DetectNewImageFeatures.cpp:172-419).  Sequential by construction (one libc rand()
// draw per candidate tried, a mask that changes after every accepted feature), a few dozen iterations per frame: it
// stays on the host and works on the new-feature mask the device built (ekfb_get_new_feature_mask).
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../include/EKF.h"

namespace {
struct Zone {   // ZoneInfo, DetectNewImageFeatures.cpp:50-69
    std::vector<int> candidates;   // indices into the kept keypoints
    int left, count, id;
};

int zoneCompare(const void* a, const void* b)   // sortCompare :73-92
{
    const Zone* z1 = *(Zone* const*)a;
    const Zone* z2 = *(Zone* const*)b;
    return z1->count < z2->count ? -1 : z1->count == z2->count ? 0 : 1;
}

int pointZone(double x, double y, int zoneWidth, int zoneHeight, int width)   // getPointZone :96-101
{
    return ((int)y / zoneHeight) * (width / zoneWidth) + (int)x / zoneWidth;
}
}  // namespace

int ekfbSelectNewFeatures(int W, int H, int divideTimes, unsigned char* mask, const float* kpXY, int nKp, const double* predXY,
                          int nPred, int maxNew, EkfbDrawFn draw, void* user, int* outIdx)
{
    if (maxNew <= 0) return 0;
    // the detector's mask filter (cv::KeyPointsFilter::runByPixelsMask inside detector->detect, :343)
    std::vector<int> kept;
    for (int i = 0; i < nKp; ++i) {
        const int yy = (int)(kpXY[2 * i + 1] + 0.5f), xx = (int)(kpXY[2 * i] + 0.5f);
        if (xx < 0 || xx >= W || yy < 0 || yy >= H || mask[(size_t)yy * W + xx] == 0) continue;
        kept.push_back(i);
    }
    int nOut = 0;
    if ((int)kept.size() <= maxNew) {   // :358-371: no more keypoints than requested -> all of them, in order
        for (size_t i = 0; i < kept.size(); ++i) outIdx[nOut++] = kept[i];
        return nOut;
    }
    // searchFeaturesByZone :172-317
    const int zonesInARow = (int)exp2f((float)divideTimes);
    const int zoneWidth = W / zonesInARow, zoneHeight = H / zonesInARow, zonesCount = zonesInARow * zonesInARow;
    std::vector<Zone> zones(zonesCount);
    std::vector<Zone*> order(zonesCount);
    for (int i = 0; i < zonesCount; ++i) {
        zones[i].left = zones[i].count = 0;
        zones[i].id = i;
        order[i] = &zones[i];
    }
    for (size_t i = 0; i < kept.size(); ++i) {
        const int z = pointZone((double)kpXY[2 * kept[i]], (double)kpXY[2 * kept[i] + 1], zoneWidth, zoneHeight, W);
        if (z < 0 || z >= zonesCount) continue;   // the reference indexes out of bounds here; such keypoints cannot exist for W, H divisible by the grid
        zones[z].candidates.push_back((int)i);
        zones[z].left++;
    }
    for (int i = 0; i < nPred; ++i) {
        const int z = pointZone(predXY[2 * i], predXY[2 * i + 1], zoneWidth, zoneHeight, W);
        if (z >= 0 && z < zonesCount) zones[z].count++;
    }
    qsort(order.data(), zonesCount, sizeof(Zone*), &zoneCompare);
    size_t front = 0;   // the std::list of :208-213 as a vector with a moving head
    int zonesLeft = zonesCount;
    while (zonesLeft > 0 && maxNew > 0) {
        Zone* cur = order[front];
        if (cur->left == 0) {
            front++;
            zonesLeft--;
            continue;
        }
        const int r = rand() % cur->left;            // :236
        const int ki = kept[cur->candidates[r]];
        const double x = (double)kpXY[2 * ki], y = (double)kpXY[2 * ki + 1];
        if (mask[(size_t)(int)y * W + (int)x]) {     // :239-241 (truncation, not the detector's rounding)
            outIdx[nOut++] = ki;
            cur->count++;
            for (size_t a = front; a + 1 < order.size(); ++a) {   // :258-283: keep the list ordered by count
                if (cur->count >= order[a + 1]->count) {
                    order[a] = order[a + 1];
                    order[a + 1] = cur;
                } else {
                    break;
                }
            }
            draw(user, mask, W, H, x, y);            // :286-291: black ellipse around the accepted feature
            maxNew--;
        }
        cur->left--;                                 // :298-300: never pick the same candidate again
        cur->candidates[r] = cur->candidates[cur->left];
    }
    return nOut;
}

namespace {
struct StampCtx { const unsigned char* stamp; int R; };
void stampDraw(void* user, unsigned char* mask, int W, int H, double x, double y)
{
    const StampCtx* s = (const StampCtx*)user;
    const int cx = (int)(float)x, cy = (int)(float)y, D = 2 * s->R + 1;   // cv::Point2d -> Point2f -> Point (Gui/Draw.cpp:51)
    for (int dy = 0; dy < D; ++dy) {
        const int yy = cy + dy - s->R;
        if (yy < 0 || yy >= H) continue;
        for (int dx = 0; dx < D; ++dx) {
            const int xx = cx + dx - s->R;
            if (xx < 0 || xx >= W) continue;
            if (s->stamp[dy * D + dx] == 0) mask[(size_t)yy * W + xx] = 0;
        }
    }
}
}  // namespace

void ekfbStampEllipse(const unsigned char* stamp, int R, unsigned char* mask, int W, int H, double x, double y)
{
    StampCtx s = {stamp, R};
    stampDraw(&s, mask, W, H, x, y);
}

// C hook for the CPU-side test (tests/test_host_ekf.py): the ellipse of every accepted feature is a (2R+1)^2 stamp
extern "C" int ekfb_host_select_new_features(int W, int H, int divideTimes, unsigned char* mask, const unsigned char* stamp, int R,
                                             const float* kpXY, int nKp, const double* predXY, int nPred, int maxNew, int* outIdx)
{
    StampCtx s = {stamp, R};
    return ekfbSelectNewFeatures(W, H, divideTimes, mask, kpXY, nKp, predXY, nPred, maxNew, &stampDraw, &s, outIdx);
}
