// oracle/ref_shim/Core/landmark.h — TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's Landmark (Core/landmark.h): the members Features/matcher.cpp reads, as plain state, and the two graph
// edits it makes (Replace, AddObservation) recorded in an event list instead of performed.
#pragma once
#include <set>
#include <vector>
#include <opencv2/core.hpp>

class KeyFrame;

class Landmark {
public:
    struct Event { int kind; Landmark* a; Landmark* b; KeyFrame* kf; size_t idx; };      // kind 0: a->AddObservation(kf, idx); 1: a->Replace(b)
    static std::vector<Event>& Log() { static std::vector<Event> log; return log; }

    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservedIn.count(pKF) > 0; }
    void AddObservation(KeyFrame* pKF, size_t idx) { Event e = { 0, this, nullptr, pKF, idx }; Log().push_back(e); mObservedIn.insert(pKF); }
    void Replace(Landmark* pLM) { Event e = { 1, this, pLM, nullptr, 0 }; Log().push_back(e); }

    bool mbTrackInView = false;
    float mTrackProjX = 0.f, mTrackProjY = 0.f;

    bool mbBad = false;
    int nObs = 0;
    cv::Mat mDescriptor, mWorldPos;
    std::set<KeyFrame*> mObservedIn;
};
