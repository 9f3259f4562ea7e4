#=GENOME_DIFF	1.0
INS	20	4	edge	5	C	frequency=3.0e-01	insert_position=1
SUB	12	2,3	edge	5	1	AA	frequency=1
DEL	21	5	edge	9	1	frequency=3.0e-01
INS	22	7	edge	14	T	frequency=3.0e-01	insert_position=1
SNP	23	9	edge	22	T	frequency=3.0e-01
SNP	24	10	edge	23	A	frequency=3.0e-01
SNP	25	11	edge	24	A	frequency=3.0e-01
INS	26	14	edge	28	A	frequency=1	insert_position=1
SNP	27	15	edge	34	T	frequency=1
INS	28	16,17	edge	39	CG	frequency=1	insert_position=1
RA	1	.	edge	1	0	A	.	deleted=1	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	frequency_lower=9.4e-01	frequency_upper=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=.	major_cov=28/30	minor_base=A	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	2	.	edge	4	1	.	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	frequency_lower=9.4e-01	frequency_upper=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	3	.	edge	5	0	T	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	frequency_lower=9.4e-01	frequency_upper=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=12/14	minor_base=T	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	4	.	edge	5	1	.	C	consensus_reject=FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	frequency_lower=2.0e-01	frequency_upper=4.0e-01	ks_quality_p_value=5.00000e-01	major_base=.	major_cov=20/22	minor_base=C	minor_cov=9/9	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=40.0	total_cov=30/30
RA	5	.	edge	9	0	G	.	consensus_reject=FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	frequency_lower=2.0e-01	frequency_upper=4.0e-01	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=20/22	minor_base=.	minor_cov=9/9	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=40.0	total_cov=30/30
RA	6	.	edge	12	0	T	G	consensus_reject=FREQUENCY_CUTOFF	fisher_strand_p_value=1.00000e-02	frequency=3.0e-01	frequency_lower=2.0e-01	frequency_upper=4.0e-01	ks_quality_p_value=1.00000e-02	major_base=T	major_cov=20/22	minor_base=G	minor_cov=9/9	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	reject=FISHER_STRAND	score=40.0	total_cov=30/30
RA	7	.	edge	14	1	.	T	consensus_reject=FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	frequency_lower=2.0e-01	frequency_upper=4.0e-01	ks_quality_p_value=5.00000e-01	major_base=.	major_cov=20/22	minor_base=T	minor_cov=9/9	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=40.0	total_cov=30/30
RA	8	.	edge	21	0	A	C	consensus_reject=FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	frequency_lower=2.0e-01	frequency_upper=4.0e-01	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=20/22	minor_base=C	minor_cov=9/1	new_cov=9/1	prediction=polymorphism	ref_cov=20/22	reject=VARIANT_STRAND_COVERAGE	score=40.0	total_cov=30/30
RA	9	.	edge	22	0	C	T	consensus_reject=EXISTING,FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	frequency_lower=2.0e-01	frequency_upper=4.0e-01	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=20/22	minor_base=T	minor_cov=9/9	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=40.0	total_cov=30/30	user_defined=1
RA	10	.	edge	23	0	G	A	consensus_reject=FREQUENCY_CUTOFF	consensus_score=5.0	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	frequency_lower=2.0e-01	frequency_upper=4.0e-01	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=20/22	minor_base=A	minor_cov=9/9	new_cov=9/9	polymorphism_score=30.0	prediction=polymorphism	ref_cov=20/22	total_cov=10/12
RA	11	.	edge	24	0	C	A	consensus_reject=FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	frequency_lower=2.0e-01	frequency_upper=4.0e-01	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=20/22	minor_base=A	minor_cov=3/3	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=NA	total_cov=30/30
RA	13	.	edge	26	0	C	T	consensus_reject=SCORE_CUTOFF,FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=2.0e-02	frequency_lower=1.0e-02	frequency_upper=4.0e-02	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=29/29	minor_base=T	minor_cov=1/1	new_cov=1/1	prediction=polymorphism	ref_cov=29/29	reject=FREQUENCY_CUTOFF,VARIANT_STRAND_COVERAGE	score=3.0	total_cov=30/30
RA	14	.	edge	28	1	.	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	frequency_lower=9.4e-01	frequency_upper=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	15	.	edge	34	0	A	T	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	frequency_lower=9.4e-01	frequency_upper=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=T	major_cov=28/30	minor_base=A	minor_cov=1/1	new_cov=28/30	note=a=b	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	16	.	edge	39	1	.	C	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	frequency_lower=9.4e-01	frequency_upper=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	17	.	edge	39	2	.	G	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	frequency_lower=9.4e-01	frequency_upper=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
MC	18	.	edge	1	2	0	0	left_inside_cov=0	left_outside_cov=NA	right_inside_cov=0	right_outside_cov=5
UN	19	.	edge	1	2
