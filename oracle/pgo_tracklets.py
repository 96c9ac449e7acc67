"""TEST INFRASTRUCTURE — CPU restatement of reconstruction::Tracklets (point_track.h:541-712), statement by statement,
with the reference's containers mapped onto their Python equivalents (std::unordered_map with operator[] -> dict with
setdefault, std::vector -> list).  Only tests/ may import this.  Parity unpinned: the reference has no tests or fixtures
for Tracklets and cannot be compiled here; what pins this file is its closeness to the source."""


class Tracklets:
    def __init__(self, viewNumber_):
        self.pointPairNumber = 0          # :552, :565
        self.pointPairs = {}              # (view, point) -> number; operator[] default-constructs 0
        self.tmpTracks = []               # :556
        self.tmpViewToTracks = {}         # :557
        self.tmpPairToTracks = {}         # :558

    def getCorrespondences(self, viewIdSource_, viewIdDestination_, maximumCorrespondenceNumber_):  # :575-636
        matches_ = []
        if viewIdSource_ not in self.tmpViewToTracks:        # :583-588
            return matches_
        if viewIdDestination_ not in self.tmpViewToTracks:   # :590-596
            return matches_
        tracksSrc = self.tmpViewToTracks[viewIdSource_]
        tracksDst = self.tmpViewToTracks[viewIdDestination_]
        trackMap = set(tracksSrc)                            # :602-606
        for trackIdx in tracksDst:                           # :608
            if trackIdx in trackMap:
                pair = [0, 0, 0.0]                           # std::tuple<size_t, size_t, double> pair;  (value-initialised)
                cnt = 0
                for p in self.tmpTracks[trackIdx]:           # :614-627
                    if p[0] == viewIdSource_:
                        pair[0] = p[1]
                        cnt += 1
                    elif p[0] == viewIdDestination_:
                        pair[1] = p[1]
                        cnt += 1
                    if cnt == 2:
                        break
                matches_.append(tuple(pair))                 # :629
                if len(matches_) > maximumCorrespondenceNumber_:  # :631-632
                    break
        return matches_

    def add(self, imageIdxSource_, imageIdxDestination_, matches_, inlierMask_):  # :638-712
        for pointIdx in range(len(matches_)):
            if not inlierMask_[pointIdx]:                    # :650-651
                continue
            pairSource = (imageIdxSource_, int(matches_[pointIdx][0]))
            pairDestination = (imageIdxDestination_, int(matches_[pointIdx][1]))
            it = self.pointPairs.setdefault(pairSource, 0)   # :655-657
            if it == 0:
                it = self.pointPairs[pairSource] = self.pointPairNumber
                self.pointPairNumber += 1
            jt = self.pointPairs.setdefault(pairDestination, 0)  # :659-661
            if jt == 0:
                jt = self.pointPairs[pairDestination] = self.pointPairNumber
                self.pointPairNumber += 1
            tracksSource = self.tmpPairToTracks.setdefault(it, [])        # :663
            tracksDestination = self.tmpPairToTracks.setdefault(jt, [])   # :664
            trackNumDestination = len(tracksDestination)                   # :665
            added = False
            for trackIdx in list(tracksSource) if tracksSource is tracksDestination else tracksSource:  # :668-679
                track = self.tmpTracks[trackIdx]
                if pairDestination in track:
                    continue
                self.tmpViewToTracks.setdefault(imageIdxDestination_, []).append(trackIdx)
                track.append(pairDestination)
                tracksDestination.append(trackIdx)
                added = True
            for trackIdxIdx in range(trackNumDestination):   # :681-693
                trackIdx = tracksDestination[trackIdxIdx]
                track = self.tmpTracks[trackIdx]
                if pairSource in track:
                    continue
                self.tmpViewToTracks.setdefault(imageIdxSource_, []).append(trackIdx)
                track.append(pairSource)
                tracksSource.append(trackIdx)
                added = True
            if not added:                                    # :695-703
                idx = len(self.tmpTracks)
                self.tmpTracks.append([pairSource, pairDestination])
                self.tmpViewToTracks.setdefault(imageIdxSource_, []).append(idx)
                self.tmpViewToTracks.setdefault(imageIdxDestination_, []).append(idx)
                tracksSource.append(idx)
                tracksDestination.append(idx)
