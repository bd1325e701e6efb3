// include/nuslam_b200/slam_library.hpp -- C++ facade (B = 1) with the reference's class and function signatures over the C ABI.
//
// What the existing callers (nuslam/src/slam.cpp:81,157,269,291,296,318; nuslam/src/landmarks.cpp:63,86,89) would compile against
// instead of nuslam/include/nuslam/{slam_library,circle_fit_library}.hpp:
//   slam_library::cartesian2polar, slam_library::ExtendedKalman            (slam_library.hpp:18-108)
//   circle_fit::clusterPoints, circle_fit::circleFit, circle_fit::classifyCluster   (circle_fit_library.hpp:18-28)
//   rigid2d::Twist2D (field order dth, dx, dy; rigid2d.hpp:150-155), rigid2d::normalize_angle (rigid2d.cpp:9-13)
// Value semantics, 1-based landmark ids, column-major matrices, exceptions where the reference throws
// (std::logic_error for Armadillo's bounds check, std::runtime_error for a singular inv()).
// With -DNUSLAM_B200_USE_ARMADILLO the facade uses arma::colvec / arma::mat; otherwise the two small value types below
// (same element access syntax). Every arithmetic operation runs in libnuslam_b200.so on an sm_100 device: no CPU path.
#ifndef NUSLAM_B200_SLAM_LIBRARY_HPP
#define NUSLAM_B200_SLAM_LIBRARY_HPP

#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../nuslam_b200.h"

#ifdef NUSLAM_B200_USE_ARMADILLO
#include <armadillo>
#endif

#ifdef NUSLAM_B200_USE_ROS
// the message types the landmarks node passes around (circle_fit_library.hpp:8-9)
#include <geometry_msgs/Point.h>
#include <visualization_msgs/Marker.h>
#endif

#ifdef NUSLAM_B200_USE_RIGID2D
// inside the reference's workspace the geometry value types stay rigid2d's own (slam.cpp also uses Vector2D, Transform2D and
// DiffDrive from it): only slam_library / circle_fit_library are replaced
#include "rigid2d/rigid2d.hpp"
#else
namespace rigid2d
{
struct Twist2D   // rigid2d.hpp:150-155: that field order
{
    double dth = 0.0, dx = 0.0, dy = 0.0;
};
inline double normalize_angle(double rad)   // rigid2d.cpp:9-13, evaluated on the device
{
    double out = 0.0;
    if (nuslam_normalize_angle(&rad, &out, 1, NUSLAM_HOST, 0, nullptr) != NUSLAM_OK) throw std::runtime_error(nuslam_last_error());
    return out;
}
}   // namespace rigid2d
#endif

namespace slam_library
{
#ifdef NUSLAM_B200_USE_ARMADILLO
using colvec = arma::colvec;
using mat = arma::mat;
inline double * data_of(colvec & v) { return v.memptr(); }
inline const double * data_of(const colvec & v) { return v.memptr(); }
inline double * data_of(mat & m) { return m.memptr(); }
inline const double * data_of(const mat & m) { return m.memptr(); }
#else
// minimal stand-ins for arma::colvec / arma::mat: column-major storage, bounds-checked operator() like Armadillo's
struct colvec
{
    std::vector<double> v;
    size_t n_elem = 0;
    colvec() = default;
    explicit colvec(size_t n) : v(n, 0.0), n_elem(n) {}
    double & operator()(size_t i)
    {
        if (i >= n_elem) throw std::logic_error("Mat::operator(): index out of bounds");
        return v[i];
    }
    const double & operator()(size_t i) const
    {
        if (i >= n_elem) throw std::logic_error("Mat::operator(): index out of bounds");
        return v[i];
    }
    double * memptr() { return v.data(); }
    const double * memptr() const { return v.data(); }
};
struct mat
{
    std::vector<double> v;
    size_t n_rows = 0, n_cols = 0, n_elem = 0;
    mat() = default;
    mat(size_t r, size_t c) : v(r * c, 0.0), n_rows(r), n_cols(c), n_elem(r * c) {}
    double & operator()(size_t i, size_t j)
    {
        if (i >= n_rows || j >= n_cols) throw std::logic_error("Mat::operator(): index out of bounds");
        return v[i + j * n_rows];
    }
    const double & operator()(size_t i, size_t j) const
    {
        if (i >= n_rows || j >= n_cols) throw std::logic_error("Mat::operator(): index out of bounds");
        return v[i + j * n_rows];
    }
    double * memptr() { return v.data(); }
    const double * memptr() const { return v.data(); }
};
inline double * data_of(colvec & v) { return v.memptr(); }
inline const double * data_of(const colvec & v) { return v.memptr(); }
inline double * data_of(mat & m) { return m.memptr(); }
inline const double * data_of(const mat & m) { return m.memptr(); }
#endif

namespace detail
{
inline void check(int rc)
{
    if (rc != NUSLAM_OK) throw std::runtime_error(std::string("nuslam_b200: ") + nuslam_last_error());
}
struct HandleDeleter
{
    void operator()(nuslam_ekf * h) const { nuslam_ekf_destroy(h); }
};
using Handle = std::unique_ptr<nuslam_ekf, HandleDeleter>;
}   // namespace detail

/// slam_library::cartesian2polar (slam_library.cpp:16-22)
inline colvec cartesian2polar(double x, double y)
{
    const double xy[2] = {x, y};
    colvec rb(2);
    detail::check(nuslam_cartesian2polar(xy, data_of(rb), 1, NUSLAM_HOST, 0, nullptr));
    return rb;
}

/// slam_library::ExtendedKalman (slam_library.hpp:23-108), one filter on the device
class ExtendedKalman
{
  public:
    ExtendedKalman() = default;   // slam_library.cpp:35-37: an empty filter, assigned later (slam.cpp:81,157)

    ExtendedKalman(colvec robotState, colvec mapState, mat Q, mat R, int device = 0, int mode = NUSLAM_MODE_STRICT)
    {
        if (robotState.n_elem != 3 || (mapState.n_elem % 2) != 0 || Q.n_elem != 9 || R.n_elem != 4)
            throw std::logic_error("ExtendedKalman: robotState(3), mapState(2n), Q(3x3), R(2x2) expected");
        nuslam_ekf_default_config(&cfg_, (int32_t) (mapState.n_elem / 2));
        cfg_.mode = mode;
        for (int k = 0; k < 9; ++k) cfg_.Q[k] = data_of(Q)[k];
        for (int k = 0; k < 4; ++k) cfg_.R[k] = data_of(R)[k];
        device_ = device;
        open();
        detail::check(nuslam_ekf_init(h_.get(), data_of(robotState), mapState.n_elem ? data_of(mapState) : nullptr, NUSLAM_HOST));
        pull();
    }

    ExtendedKalman(const ExtendedKalman & o) { *this = o; }
    ExtendedKalman & operator=(const ExtendedKalman & o)   // value semantics: deep copy of the device state
    {
        if (this == &o) return *this;
        cfg_ = o.cfg_;
        device_ = o.device_;
        x_ = o.x_;
        sigma_ = o.sigma_;
        seen_ = o.seen_;
        h_.reset();
        if (o.h_)
        {
            open();
            int32_t status = 0;
            detail::check(nuslam_ekf_get_state(o.h_.get(), nullptr, nullptr, nullptr, &status, NUSLAM_HOST));
            detail::check(nuslam_ekf_set_state(h_.get(), data_of(x_), data_of(sigma_), &seen_, &status, NUSLAM_HOST));
        }
        return *this;
    }
    ExtendedKalman(ExtendedKalman &&) = default;
    ExtendedKalman & operator=(ExtendedKalman &&) = default;

    /// slam_library.cpp:65-69
    void predict(const rigid2d::Twist2D & tw)
    {
        const double t[3] = {tw.dth, tw.dx, tw.dy};
        detail::check(nuslam_ekf_predict(need(), t, NUSLAM_HOST));
        pull();
    }

    /// slam_library.cpp:150-160; evaluated at `state`, not at the filter's own state (the reference takes it by value)
    colvec computeTheoreticalMeasurement(int j, colvec state) const
    {
        colvec zhat(2);
        model_at(j, state, data_of(zhat), nullptr);
        return zhat;
    }

    /// slam_library.cpp:162-186
    mat linearizedMeasurementModel(int j, colvec state) const
    {
        mat H(2, state.n_elem);
        model_at(j, state, nullptr, data_of(H));
        return H;
    }

    /// slam_library.cpp:188-253: k >= 1, or -1 (ambiguous); throws where the reference's bounds check throws (full map)
    int associateLandmark(colvec z)
    {
        if (z.n_elem != 2) throw std::logic_error("associateLandmark: z must have 2 elements");
        int32_t id = 0;
        detail::check(nuslam_ekf_associate(need(), data_of(z), &id, NUSLAM_HOST));
        pull();
        if (id == NUSLAM_ID_EXCEPTION) throw_for_status();
        return id;
    }

    /// slam_library.cpp:255-261
    void initializeLandmark(colvec z, int id)
    {
        const int32_t i = id;
        if (id < 1 || id > cfg_.n_landmarks) throw std::logic_error("Mat::operator(): index out of bounds");
        detail::check(nuslam_ekf_initialize_landmark(need(), data_of(z), &i, NUSLAM_HOST));
        pull();
    }

    /// slam_library.cpp:263-282 (`tw` is unused by the reference too)
    void update(const rigid2d::Twist2D &, colvec z, int id)
    {
        const int32_t i = id;
        if (id < 1 || id > cfg_.n_landmarks) throw std::logic_error("Mat::operator(): index out of bounds");
        detail::check(nuslam_ekf_update(need(), data_of(z), &i, NUSLAM_HOST));
        pull();
        throw_for_status();
    }

    const colvec & getStateVector() const { return x_; }       // slam_library.cpp:284-287
    const mat & getCovariance() const { return sigma_; }       // :289-292
    const int & getSeenLandmarks() const { return seen_; }     // :294-297

  private:
    void open()
    {
        nuslam_ekf * raw = nullptr;
        detail::check(nuslam_ekf_create(&cfg_, 1, device_, nullptr, &raw));
        h_.reset(raw);
        const size_t len = 3 + 2 * (size_t) cfg_.n_landmarks;
        if (x_.n_elem != len)
        {
            x_ = colvec(len);
            sigma_ = mat(len, len);
        }
    }
    nuslam_ekf * need() const
    {
        if (!h_) throw std::logic_error("ExtendedKalman: default-constructed filter used before assignment");
        return h_.get();
    }
    void pull()
    {
        int32_t seen = 0;
        detail::check(nuslam_ekf_get_state(h_.get(), data_of(x_), data_of(sigma_), &seen, nullptr, NUSLAM_HOST));
        seen_ = seen;
    }
    void throw_for_status() const
    {
        int32_t status = 0;
        detail::check(nuslam_ekf_get_state(h_.get(), nullptr, nullptr, nullptr, &status, NUSLAM_HOST));
        if (status & (NUSLAM_FILTER_MAP_FULL | NUSLAM_FILTER_BAD_ID)) throw std::logic_error("Mat::operator(): index out of bounds");
        if (status & NUSLAM_FILTER_SINGULAR) throw std::runtime_error("inv(): matrix is singular");
    }
    void model_at(int j, const colvec & state, double * zhat, double * H) const
    {
        if (state.n_elem != x_.n_elem) throw std::logic_error("state vector length mismatch");
        if (j < 1 || j > cfg_.n_landmarks) throw std::logic_error("Mat::operator(): index out of bounds");
        // a scratch filter holds `state`; the measurement model kernels read a filter's state
        if (!scratch_)
        {
            nuslam_ekf * raw = nullptr;
            detail::check(nuslam_ekf_create(&cfg_, 1, device_, nullptr, &raw));
            scratch_.reset(raw);
        }
        detail::check(nuslam_ekf_set_state(scratch_.get(), data_of(state), nullptr, nullptr, nullptr, NUSLAM_HOST));
        const int32_t jj = j;
        detail::check(nuslam_ekf_measurement_model(scratch_.get(), &jj, zhat, H, NUSLAM_HOST));
    }

    nuslam_ekf_config cfg_{};
    int device_ = 0;
    detail::Handle h_;
    mutable detail::Handle scratch_;
    colvec x_;
    mat sigma_;
    int seen_ = 0;
};
}   // namespace slam_library

namespace circle_fit
{
#ifndef NUSLAM_B200_USE_ROS
// stand-ins for geometry_msgs::Point and the fields of visualization_msgs::Marker the callers read (landmarks.cpp:91-105, slam.cpp:282-283)
struct Point
{
    double x = 0.0, y = 0.0, z = 0.0;
};
struct Marker
{
    int id = 0;
    struct
    {
        Point position;
    } pose;
    struct
    {
        double x = 0.0, y = 0.0, z = 0.0;
    } scale;
};
#else
using Point = geometry_msgs::Point;
using Marker = visualization_msgs::Marker;
#endif

/// circle_fit::clusterPoints (circle_fit_library.cpp:136-206). The clustering runs on the device; the facade rebuilds the
/// reference's return type (points r (cos, sin)(deg2rad(i)) in stored order, the wrap point last in cluster 0).
inline std::vector<std::vector<Point>> clusterPoints(std::vector<float> ranges, double minRange, double maxRange)
{
    if (ranges.size() < NUSLAM_SCAN_BEAMS) throw std::logic_error("clusterPoints: the reference reads exactly 360 beams");
    std::vector<int16_t> cob(NUSLAM_SCAN_BEAMS);
    int32_t ncl = 0, nci = 0;
    double circles[4];
    slam_library::detail::check(nuslam_scan_detect(ranges.data(), 1, minRange, maxRange, cob.data(), &ncl, &nci, circles, 1, NUSLAM_HOST, 0, nullptr));
    if (nci == NUSLAM_SCAN_UB) throw std::logic_error("clusterPoints: clusters[0] of an empty vector (undefined behaviour in the reference)");
    std::vector<std::vector<Point>> clusters((size_t) ncl);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < NUSLAM_SCAN_BEAMS; ++i)
        if (cob[i] >= 0)
        {
            Point p;
            p.x = ranges[i] * std::cos((pi / (double) 180) * i);
            p.y = ranges[i] * std::sin((pi / (double) 180) * i);
            clusters[(size_t) cob[i]].push_back(p);
        }
    return clusters;
}

namespace detail
{
inline void fit(const std::vector<Point> & data, int32_t & is_circle, double * fit4)
{
    std::vector<double> px(data.size()), py(data.size());
    for (size_t k = 0; k < data.size(); ++k)
    {
        px[k] = data[k].x;
        py[k] = data[k].y;
    }
    const int32_t off[2] = {0, (int32_t) data.size()};
    slam_library::detail::check(nuslam_classify_and_fit(px.data(), py.data(), off, 1, &is_circle, fit4, NUSLAM_HOST, 0, nullptr));
}
}   // namespace detail

/// circle_fit::circleFit (circle_fit_library.cpp:15-134): centre in pose.position, scale.x = scale.y = 2R, id = -1 for < 4 points
inline Marker circleFit(std::vector<Point> data)
{
    int32_t is_circle = 0;
    double f[4] = {0, 0, 0, 0};
    detail::fit(data, is_circle, f);
    Marker m;
    m.id = (int) f[0];
    if (m.id < 0) return m;
    m.pose.position.x = f[1];
    m.pose.position.y = f[2];
    m.pose.position.z = 0.25;
    m.scale.x = 2 * f[3];
    m.scale.y = 2 * f[3];
    m.scale.z = 0.5;
    return m;
}

/// circle_fit::classifyCluster (circle_fit_library.cpp:208-250)
inline bool classifyCluster(std::vector<Point> cluster)
{
    int32_t is_circle = 0;
    double f[4];
    detail::fit(cluster, is_circle, f);
    return is_circle != 0;
}
}   // namespace circle_fit

#endif
