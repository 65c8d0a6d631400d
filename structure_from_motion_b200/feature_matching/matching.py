"""Mirror of the ``Match`` value type of lib/feature_matching/matching.py:15-24.

Only the dataclass crosses the hot-path boundary; the brute-force matcher itself
(matching.py:27-118) is upstream of the path and out of scope.
"""
import dataclasses
import math


@dataclasses.dataclass
class Match:
    a_index: int = -1
    b_index: int = -1
    # A lower score indicates a better match in all cases.
    match_score: float = math.inf

    def __lt__(self, other) -> bool:
        return self.match_score < other.match_score
