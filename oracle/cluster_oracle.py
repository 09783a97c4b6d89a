"""CPU ORACLE (test infrastructure) for the match-clustering stage behind Detector::match -- a pure-Python restatement of
rgbdDetector::rcd_voting (/root/reference/src/rgbdDetector.cpp:36-70), cluster_filter (72-84, intended semantics: the
upstream loop erases from the std::map it iterates), similarity_score_calc (133-145), nonMaximaSuppressionUsingIOU
(462-533) and computeIoU (535-574).  Only tests/ may import this module."""
import numpy as np


def compute_iou(r1, r2):
    """rects are (x, y, width, height); float32 arithmetic like the reference's `float` locals."""
    r1_minx, r1_maxx, r1_miny, r1_maxy = r1[0], r1[0] + r1[2] - 1, r1[1], r1[1] + r1[3] - 1
    r2_minx, r2_maxx, r2_miny, r2_maxy = r2[0], r2[0] + r2[2] - 1, r2[1], r2[1] + r2[3] - 1
    minx, maxx = max(r1_minx, r2_minx), min(r1_maxx, r2_maxx)
    miny, maxy = max(r1_miny, r2_miny), min(r1_maxy, r2_maxy)
    x_inter = (r1_minx <= minx <= r1_maxx) or (r2_minx <= minx <= r2_maxx)
    y_inter = (r1_miny <= miny <= r1_maxy) or (r2_miny <= miny <= r2_maxy)
    inter = np.float32((maxx - minx + 1) * (maxy - miny + 1)) if (x_inter and y_inter) else np.float32(0.0)
    union = np.float32(r1[2] * r1[3] + r2[2] * r2[3]) - inter
    return np.float32(inter / union)


def _trunc_div(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def cluster_matches(matches, obj_origin_dists, rects, vote_step, radius_min, radius_step, cluster_threshold=2, iou_threshold=0.4):
    """matches: structured array (x, y, template_id, class_index, similarity) in Detector::match order.
    Returns [(index(3), score, rect(4), [match indices])] in the reference's output order."""
    depth_step = np.float32(radius_step)
    bins = {}
    for i, m in enumerate(matches):
        depth = np.float32(obj_origin_dists[int(m["template_id"])])
        idx = (_trunc_div(int(m["y"]), vote_step), _trunc_div(int(m["x"]), vote_step),
               int((float(depth) - float(radius_min)) / float(depth_step)))      # (int) truncates toward zero
        bins.setdefault(idx, []).append(i)
    clusters = []
    for idx in sorted(bins):                                                      # std::map<std::vector<int>> order
        members = bins[idx]
        if len(members) <= cluster_threshold:
            continue
        score = 0.0
        for i in members:
            score += float(matches[i]["similarity"])
        score /= len(members)
        n = len(members)
        X = _trunc_div(sum(int(matches[i]["x"]) for i in members), n)
        Y = _trunc_div(sum(int(matches[i]["y"]) for i in members), n)
        Wd = _trunc_div(sum(int(rects[int(matches[i]["template_id"])][2]) for i in members), n)
        Ht = _trunc_div(sum(int(rects[int(matches[i]["template_id"])][3]) for i in members), n)
        clusters.append([idx, score, (X, Y, Wd, Ht), members, False])
    # std::sort by score descending (tests keep scores distinct: std::sort is not stable beyond 16 elements)
    clusters.sort(key=lambda c: -c[1])
    for i, ci in enumerate(clusters):
        if ci[4]:
            continue
        for cj in clusters[i + 1:]:
            if not cj[4] and float(compute_iou(ci[2], cj[2])) > iou_threshold:
                cj[4] = True
    return [(c[0], c[1], c[2], c[3]) for c in clusters if not c[4]]
