// tests/facade_main.cpp -- exercises the C++ facade exactly as nuslam/src/slam.cpp:262-319 and landmarks.cpp:84-109 use the reference
// library: prints state / covariance / ids so that tests/test_facade.py can compare with the oracle.
#include <cstdio>
#include "nuslam_b200/slam_library.hpp"

int main()
{
    using namespace slam_library;
    colvec robot(3), map(8);
    robot(0) = 0.1;
    robot(1) = -0.2;
    robot(2) = 0.3;
    mat Q(3, 3), R(2, 2);
    for (int i = 0; i < 3; ++i) Q(i, i) = 0.1;
    for (int i = 0; i < 2; ++i) R(i, i) = 0.001;
    ExtendedKalman ekf;                            // slam.cpp:81
    ekf = ExtendedKalman(robot, map, Q, R);        // slam.cpp:157
    rigid2d::Twist2D tw;
    tw.dth = 0.02;
    tw.dx = 0.007;
    const double zs[3][2] = {{1.0, 0.1}, {2.0, -1.0}, {3.0, 2.0}};
    for (int step = 0; step < 3; ++step)
    {
        ekf.predict(tw);
        const int seen0 = ekf.getSeenLandmarks();   // slam.cpp:251
        for (int k = 0; k < 3; ++k)
        {
            colvec z(2);
            z(0) = zs[k][0];
            z(1) = zs[k][1];
            int id;
            try
            {
                id = ekf.associateLandmark(z);
            }
            catch (const std::logic_error & e)
            {
                printf("EXC %d %d %s\n", step, k, e.what());
                continue;
            }
            printf("ID %d %d %d\n", step, k, id);
            if (id > seen0) ekf.initializeLandmark(z, id);
            else if (id < 0) continue;
            ekf.update(tw, z, id);
        }
    }
    const colvec & x = ekf.getStateVector();
    printf("X");
    for (size_t i = 0; i < x.n_elem; ++i) printf(" %.17g", x(i));
    printf("\nS");
    const mat & S = ekf.getCovariance();
    for (size_t j = 0; j < S.n_cols; ++j)
        for (size_t i = 0; i < S.n_rows; ++i) printf(" %.17g", S(i, j));
    printf("\nSEEN %d\n", ekf.getSeenLandmarks());
    colvec rb = cartesian2polar(3.0, -4.0);
    printf("C2P %.17g %.17g\n", rb(0), rb(1));
    colvec zh = ekf.computeTheoreticalMeasurement(1, x);
    printf("ZHAT %.17g %.17g\n", zh(0), zh(1));
    // circle path, circle_tests.cpp:15-40
    std::vector<circle_fit::Point> data(6);
    const double pts[6][2] = {{1, 7}, {2, 6}, {5, 8}, {7, 7}, {9, 5}, {3, 7}};
    for (int k = 0; k < 6; ++k)
    {
        data[k].x = pts[k][0];
        data[k].y = pts[k][1];
    }
    circle_fit::Marker mk = circle_fit::circleFit(data);
    printf("FIT %d %.17g %.17g %.17g\n", mk.id, mk.pose.position.x, mk.pose.position.y, mk.scale.x);
    // a fourth landmark fills the map (n = 4); the next new one makes the reference's bounds check throw
    // (slam_library.cpp:206): the facade throws std::logic_error too
    {
        colvec z(2);
        z(0) = 0.3;
        z(1) = -2.5;
        const int id = ekf.associateLandmark(z);
        printf("FOURTH %d\n", id);
        ekf.initializeLandmark(z, id);
        ekf.update(tw, z, id);
        z(0) = 5.0;
        z(1) = 3.0;
        try
        {
            const int id5 = ekf.associateLandmark(z);
            printf("FULL id %d\n", id5);
        }
        catch (const std::logic_error & e)
        {
            printf("FULL EXC %s\n", e.what());
        }
    }
    std::vector<float> ranges(360, 2.0f);
    for (int i = 40; i < 47; ++i) ranges[i] = 0.5f + 0.002f * (i - 43) * (i - 43);
    auto cl = circle_fit::clusterPoints(ranges, 0.05, 1.0);
    printf("CLUSTERS %zu", cl.size());
    for (auto & c : cl) printf(" %zu", c.size());
    printf("\nCIRCLE %d\n", cl.empty() ? -1 : (int) circle_fit::classifyCluster(cl[0]));
    return 0;
}
