#=GENOME_DIFF	1.0
RA	1	.	edge	1	0	A	.	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=.	major_cov=28/30	minor_base=A	minor_cov=1/1	new_cov=28/30	ref_cov=1/1	score=40.0	total_cov=30/30
RA	2	.	edge	4	1	.	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	ref_cov=1/1	score=40.0	total_cov=30/30
RA	3	.	edge	5	0	T	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=12/14	minor_base=T	minor_cov=1/1	new_cov=28/30	ref_cov=1/1	score=40.0	total_cov=30/30
RA	4	.	edge	5	1	.	C	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=.	major_cov=20/22	minor_base=C	minor_cov=9/9	new_cov=9/9	ref_cov=20/22	score=40.0	total_cov=1/0
RA	5	.	edge	9	0	G	.	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=20/22	minor_base=.	minor_cov=9/9	new_cov=9/9	ref_cov=20/22	score=40.0	total_cov=30/30
RA	6	.	edge	12	0	T	G	fisher_strand_p_value=1.00000e-02	frequency=3.0e-01	ks_quality_p_value=1.00000e-02	major_base=T	major_cov=20/22	minor_base=G	minor_cov=9/9	new_cov=9/9	ref_cov=20/22	score=40.0	total_cov=0/0
RA	7	.	edge	14	1	.	T	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=.	major_cov=20/22	minor_base=T	minor_cov=9/9	new_cov=9/9	ref_cov=20/22	score=40.0	total_cov=2/1
RA	8	.	edge	21	0	A	C	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=20/22	minor_base=C	minor_cov=9/1	new_cov=9/1	ref_cov=20/22	score=40.0	total_cov=400/380
RA	9	.	edge	22	0	C	T	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=20/22	minor_base=T	minor_cov=9/9	new_cov=9/9	ref_cov=20/22	reject=EXISTING	score=40.0	total_cov=30/30	user_defined=1
RA	10	.	edge	23	0	G	A	consensus_score=5.0	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=20/22	minor_base=A	minor_cov=9/9	new_cov=9/9	polymorphism_score=30.0	ref_cov=20/22	total_cov=10/12
RA	11	.	edge	24	0	C	A	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=20/22	minor_base=A	minor_cov=3/3	new_cov=9/9	ref_cov=20/22	score=NA	total_cov=30/30
RA	12	.	edge	25	0	C	G	fisher_strand_p_value=5.00000e-01	frequency=0.0e+00	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=30/30	minor_base=G	minor_cov=0/0	new_cov=0/0	ref_cov=30/30	score=40.0	total_cov=30/30
RA	13	.	edge	26	0	C	T	fisher_strand_p_value=5.00000e-01	frequency=2.0e-02	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=29/29	minor_base=T	minor_cov=1/1	new_cov=1/1	ref_cov=29/29	score=3.0	total_cov=30/30
RA	14	.	edge	28	1	.	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	ref_cov=1/1	score=40.0	total_cov=30/30
RA	15	.	edge	34	0	A	T	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=T	major_cov=28/30	minor_base=A	minor_cov=1/1	new_cov=28/30	note=a=b	ref_cov=1/1	score=40.0	total_cov=30/30
RA	16	.	edge	39	1	.	C	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	ref_cov=1/1	score=40.0	total_cov=30/30
RA	17	.	edge	39	2	.	G	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	ref_cov=1/1	score=40.0	total_cov=30/30
MC	18	.	edge	1	2	0	0	left_inside_cov=0	left_outside_cov=NA	right_inside_cov=0	right_outside_cov=5
UN	19	.	edge	1	2
